"""Is the tet pass DRAM-latency bound?  Same mesh density as C4 on a 256^3 float volume (67 MB, fits the 126 MB L2):
time with L2 flushed before each launch vs. left warm."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
syn = pkg.synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
edge = 512 / 67.0
vol = syn.shells_phantom(n, levels=(0.1, 0.5, 0.9), radii=(0.40, 0.24), sigma=0.02, seed=1234, dtype=np.float32)
img = pkg.Image3D(vol)
seg = pkg.SegmentFunction(img, 3, 0.01)


def tet_mean(pos, tets, base):
    seg.set_snapshot({"pos": pos, "tet_nodes": tets, "tet_label": base, "face_nodes": np.zeros((0, 3), np.uint32),
                      "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)})
    return seg.tet_integrate(include_bound=True)["mean"]


snap = syn.build_case(vol, edge, 3, (0.3, 0.7), tet_mean, jitter=0.2, seed=5)
seg.set_snapshot(snap)
seg.enable_timers(True)
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
print("n=%d tets=%d" % (n, seg.n_tets))
for variant, lanes, label in ((0, 8, "v1 G=8"), (1, 0, "tpt"), (3, 0, "tpt tex"), (8 | 4, 8, "g2 G=8 f=1")):
    seg.set_tet_variant(variant)
    seg.set_lanes_per_tet(lanes)
    for mode in ("cold", "warm"):
        ts = []
        for i in range(6):
            if mode == "cold":
                flush.add_(1)
            torch.cuda.synchronize()
            seg.tet_integrate(per_tet=False)
            ts.append(seg.timer_ms("tet"))
        ns = seg.last_sample_count()
        print("%-12s %s: best %.4f ms  (%.1f Gsamples/s)" % (label, mode, min(ts), ns / min(ts) / 1e6))
