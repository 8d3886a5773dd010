"""Times every implementation variant of the per-tet integration kernel on a BASELINE config and checks that they agree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
syn = pkg.synthetic


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    jitter = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
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

    snap = syn.build_case(vol, edge, P, cfg["thresholds"], tet_mean, jitter=jitter, seed=5)
    seg.set_snapshot(snap)
    seg.enable_timers(True)
    print("%s: %d tets, %d faces, jitter %.2f" % (name, seg.n_tets, seg.n_faces, jitter))
    ref = None
    import torch
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    cases = [(128 | 16384, 0, "w2 default"), (131072 | 128 | 16384, 0, "staged")]
    for variant, lanes, label in cases:
        seg.set_tet_variant(variant)
        seg.set_lanes_per_tet(lanes)
        ts = []
        for i in range(6):
            flush.add_(1)
            torch.cuda.synchronize()
            o = seg.tet_integrate(per_tet=False)
            ts.append(seg.timer_ms("tet"))
        full = seg.tet_integrate(per_tet=True)
        if ref is None:
            ref = full
            err = 0.0
        else:
            lab = snap["tet_label"] != 0
            err = float(np.max(np.abs(full["total"][lab] - ref["total"][lab]) / np.maximum(np.abs(ref["total"][lab]), 1e-30)))
            assert np.array_equal(full["vol"], ref["vol"])
        print("variant %d (%-15s): best %.3f ms  median %.3f ms  max rel per-tet diff vs v0 %.2e  means %s" % (
            variant, label, min(ts), float(np.median(ts)), err, o["phase_mean"]))


if __name__ == "__main__":
    main()
