#!/usr/bin/env python3
"""Writes tests/golden/deformed_e12.dsc with the REFERENCE's own writer (is_mesh::export_tet_mesh through oracle/_ref/dsc_ref
--save-dsc) -- the square_round_2 mesh at edge 12 after one full segment() iteration -- and tests/golden/deformed_e12_dsc.npz with
the snapshot the same run dumped (raw keys compacted the way extract_tet_mesh compacts them).  Container only (needs /root/reference).
usage: python tests/golden/make_dsc_fixture.py"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402

vol = make_golden.png_stack("square_round_2")
with tempfile.TemporaryDirectory() as tmp:
    vp = os.path.join(tmp, "vol.raw")
    vol.tofile(vp)
    out = os.path.join(tmp, "out")
    dsc = os.path.join(HERE, "deformed_e12.dsc")
    subprocess.run([make_golden.REF_BIN, "--mode", "dump", "--volume", vp, "--vol-dtype", "u8", "--dims", "100", "100", "100", "--edge", "12", "--nb-phase", "3",
                    "--alpha", "0.001", "--labels", "otsu", "--iters", "1", "--out", out, "--save-dsc", dsc], check=True, capture_output=True)
    rd = lambda n, dt: np.fromfile(os.path.join(out, n), dtype=dt)
    pos = rd("pos.f64", np.float64).reshape(-1, 3)
    tets = rd("tet_nodes.u32", np.uint32).reshape(-1, 4)
    lab = rd("tet_label.i32", np.int32)
    meta = json.load(open(os.path.join(out, "meta.json")))
np.savez_compressed(os.path.join(HERE, "deformed_e12_dsc.npz"), pos=pos, tet_nodes=tets, tet_label=lab)
print("wrote", dsc, os.path.getsize(dsc), "bytes;", len(tets), "tets", len(pos), "node slots")
