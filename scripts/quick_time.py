"""Quick device timings of every kernel class on a BASELINE config (development helper)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
syn = pkg.synthetic


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    cfg = syn.CONFIGS[name]
    t0 = time.time()
    vol = syn.make_volume(name)
    print("volume", vol.shape, vol.dtype, "%.1fs" % (time.time() - t0), flush=True)
    img = pkg.Image3D(vol)
    P = cfg["nb_phase"]
    seg = pkg.SegmentFunction(img, P, 0.01)
    seg.enable_timers(True)
    edge = syn.edge_for_cells(cfg["n"], cfg["cells"])

    def tet_mean(pos, tets, base):
        snap0 = {"pos": pos, "tet_nodes": tets, "tet_label": base, "face_nodes": np.zeros((0, 3), np.uint32),
                 "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)}
        seg.set_snapshot(snap0)
        return seg.tet_integrate(include_bound=True)["mean"]

    t0 = time.time()
    snap = syn.build_case(vol, edge, P, cfg["thresholds"], tet_mean, jitter=float(os.environ.get("JITTER", "0.2")), seed=5)
    print("mesh: %d nodes %d tets %d interface faces, edge %.3f, %.1fs" % (len(snap["pos"]), len(snap["tet_nodes"]), len(snap["face_nodes"]), edge,
                                                                        time.time() - t0), flush=True)
    print("labels", np.bincount(snap["tet_label"]))
    seg.set_snapshot(snap)
    for lanes in (8, 32):
        seg.set_lanes_per_tet(lanes)
        ts = []
        for i in range(6):
            o = seg.tet_integrate(per_tet=False)
            ts.append(seg.timer_ms("tet"))
        ns = seg.last_sample_count()
        print("tet lanes=%d: ms %s  samples %d (%.1f/tet)  mean %s" % (lanes, np.round(ts, 3), ns, ns / max(1, (snap["tet_label"] != 0).sum()), o["phase_mean"]))
    seg.set_lanes_per_tet(0)
    img.build_sum_table()
    print("sumtable ms", seg.timer_ms("sumtable"))
    for i in range(3):
        t0 = time.time()
        m = seg.update_average_intensity()
        print("zray ms (events, incl. host sync inside)", round(seg.timer_ms("zray"), 3), "wall", round((time.time() - t0) * 1e3, 3), m, seg._phase_volume)
    for mode, nm in ((pkg.FORCE_GATHER, "gather"), (pkg.FORCE_ATOMIC, "atomic")):
        ts = []
        for i in range(5):
            F = seg.compute_external_force(mode=mode, mean=m)
            ts.append(seg.timer_ms("force"))
        print("force %s ms %s gauss %d |F| %.6f" % (nm, np.round(ts, 3), seg.last_gauss_count(), np.linalg.norm(F)))
    for i in range(2):
        t0 = time.time()
        d, a = seg.compute_energy(mean=m)
        print("energy", d, a, "zray-energy ms", round(seg.timer_ms("zray"), 3), "wall", round((time.time() - t0) * 1e3, 3))
    for mm, nm in ((pkg.MEAN_TET, "tet"), (pkg.MEAN_ZRAY, "zray")):
        for i in range(3):
            t0 = time.time()
            seg.image_force_eval(snap, mean_mode=mm, force_mode=pkg.FORCE_GATHER)
            wall = (time.time() - t0) * 1e3
        print("e2e image_force_eval(%s) wall ms %.3f  h2d %.3f tet %.3f force %.3f d2h %.3f" % (nm, wall, seg.timer_ms("h2d"), seg.timer_ms("tet"),
                                                                                             seg.timer_ms("force"), seg.timer_ms("d2h")))
    print("launches", seg.launch_count())


if __name__ == "__main__":
    main()
