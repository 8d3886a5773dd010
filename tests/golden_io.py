"""Loader for the reference-produced fixtures in tests/golden (see tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["c1_e6", "c1_e8_it2", "c1n_e9_j", "sph48_f32", "sh40_u16", "c1_crop_xyz"]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(bytes(g.pop("meta_json")).decode())
    return g


def snapshot_of(g):
    """The SoA snapshot (the C ABI's input) contained in a golden case."""
    return {"pos": g["pos"], "tet_nodes": g["tet_nodes"], "tet_label": g["tet_label"], "face_nodes": g["face_nodes"],
            "face_tets": g["face_tets"], "face_is_boundary": g["face_is_boundary"]}
