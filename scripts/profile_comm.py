"""The exchange kernels of a multi-GPU evaluation on ONE GPU (world = 1: every push and wait is local), for ncu: builds a BASELINE
config, takes a 1/PART share of the tets and faces like one of PART ranks would, and runs the fused evaluation a few times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

pkg = load_package()


def main():
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    part = int(os.environ.get("PART", "8"))

    def factory(vol, P):
        img = pkg.Image3D(vol)
        return img, pkg.SegmentFunction(img, P, 0.01)

    vol, cfg, edge, snap, img, seg = bench.build_workload(pkg, name, factory, 0.2)
    P = cfg["nb_phase"]
    seg.set_snapshot(snap)
    nt, nf, nn = len(snap["tet_nodes"]), len(snap["face_nodes"]), len(snap["pos"])
    seg.comm_create(0, 1, cap_rows=nn)
    r = part // 2
    seg.set_partition((nt * r // part, nt * (r + 1) // part), (nf * r // part, nf * (r + 1) // part), (0, nn))
    d_sums = torch.zeros(3 * P, dtype=torch.float64, device="cuda")
    d_mean = torch.zeros(P, dtype=torch.float64, device="cuda")
    d_F = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
    seg.enable_timers(True)
    for i in range(reps):
        seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
        seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_F.data_ptr())
        img.synchronize()
        print("tet ms", seg.timer_ms("tet"), "force ms", seg.timer_ms("force"), d_mean.cpu().numpy())
    seg.comm_error()


if __name__ == "__main__":
    main()
