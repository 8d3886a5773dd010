"""The host side of the drop-in without a GPU: dscgpu::snapshot_from_dsc (include/dscgpu_adapter.hpp) walks a live is_mesh/DSC complex
built from the unmodified reference sources (oracle/walk_timing.cpp -> oracle/_ref/dsc_walk) and must flatten it into exactly the arrays
the reference driver dumped for the same case (tests/golden/c1_e6.npz: grid mesh edge 6, Otsu initialisation)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WALK = os.path.join(ROOT, "oracle", "_ref", "dsc_walk")


def fnv1a(chunks):
    # vectorised FNV-1a is not possible (it is sequential); 1.5 MB in pure Python takes a second or two
    h = 1469598103934665603
    for c in chunks:
        for b in c.tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.skipif(not os.path.exists(WALK), reason="oracle/_ref/dsc_walk is built in the container (make -C oracle walk)")
def test_adapter_walk_flattens_the_complex_like_the_reference_driver():
    g = golden_io.load("c1_e6")
    with tempfile.TemporaryDirectory() as tmp:
        vp = os.path.join(tmp, "vol.u8")
        g["volume"].tofile(vp)
        r = subprocess.run([WALK, vp, "100", "100", "100", "6", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("walk ")]
    assert len(lines) == 2
    m = [re.search(r"(\d+) tet slots, (\d+) interface faces, (\d+) nodes \| snapshot hash ([0-9a-f]{16})", l) for l in lines]
    assert all(m)
    assert m[0].group(4) == m[1].group(4)  # the second walk reads the complex's caches: same arrays
    assert (int(m[0].group(1)), int(m[0].group(2)), int(m[0].group(3))) == (g["meta"]["n_tets"], g["meta"]["n_if"], len(g["pos"]))
    # the same bytes as the reference driver's dump (a fresh grid has no free tet slots: key-indexed = dense)
    expect = fnv1a([np.ascontiguousarray(g["pos"], np.float64), np.ascontiguousarray(g["tet_nodes"], np.uint32), np.ascontiguousarray(g["tet_label"], np.int32),
                    np.ascontiguousarray(g["face_nodes"], np.uint32), np.ascontiguousarray(g["face_tets"], np.uint32),
                    np.ascontiguousarray(g["face_is_boundary"], np.uint8), np.ascontiguousarray(g["node_flags"] & 3, np.uint8),
                    np.ascontiguousarray(g["face_key"], np.uint32)])
    assert "%016x" % expect == m[0].group(4)
