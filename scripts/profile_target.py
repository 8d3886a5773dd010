"""Short fixed workload for ncu: builds a BASELINE config and runs each kernel class a few times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
syn = pkg.synthetic


def build(name, seg):
    cfg = syn.CONFIGS[name]
    vol = syn.make_volume(name)
    return vol, cfg


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    what = sys.argv[3] if len(sys.argv) > 3 else "all"
    cfg = syn.CONFIGS[name]
    vol = syn.make_volume(name)
    img = pkg.Image3D(vol)
    P = cfg["nb_phase"]
    seg = pkg.SegmentFunction(img, P, 0.01)
    edge = syn.edge_for_cells(cfg["n"], cfg["cells"])

    def tet_mean(pos, tets, base):
        seg.set_snapshot({"pos": pos, "tet_nodes": tets, "tet_label": base, "face_nodes": np.zeros((0, 3), np.uint32),
                          "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)})
        return seg.tet_integrate(include_bound=True)["mean"]

    snap = syn.build_case(vol, edge, P, cfg["thresholds"], tet_mean, jitter=0.2, seed=5)
    seg.set_snapshot(snap)
    part = int(os.environ.get("PART", "1"))
    if part > 1:  # what one of `part` ranks sees: a contiguous 1/part of the tets (from the middle of the mesh) and of the faces
        nt, nf, nn = len(snap["tet_nodes"]), len(snap["face_nodes"]), len(snap["pos"])
        r = part // 2
        seg.set_partition((nt * r // part, nt * (r + 1) // part), (nf * r // part, nf * (r + 1) // part), (0, nn))
    seg.enable_timers(True)
    if what == "hot":  # the instantiation of the per-iteration hot path: no per-tet outputs, no second moment
        import torch
        d_sums = torch.zeros(3 * P, dtype=torch.float64, device="cuda")
        for i in range(reps):
            seg.tet_phase_sums_dev(d_sums.data_ptr())
            img.synchronize()
            print("tet ms", seg.timer_ms("tet"), d_sums.cpu().numpy())
        return
    for i in range(reps):
        if what in ("all", "tet"):
            o = seg.tet_integrate(per_tet=False)
            print("tet ms", seg.timer_ms("tet"), o["phase_mean"])
        if what in ("all", "zray"):
            m = seg.update_average_intensity()
            print("zray ms", seg.timer_ms("zray"))
        if what in ("all", "force"):
            seg.compute_external_force(mean=o["phase_mean"] if what != "force" else [0.1, 0.5, 0.9][:P])
            print("force ms", seg.timer_ms("force"))
    print("done", seg.launch_count())


if __name__ == "__main__":
    main()
