#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Runs oracle/_ref/dsc_ref (built by `make -C oracle ref` from the sources under /root/reference,
see oracle/ref_driver.cpp) on a handful of small cases and packs the flattened mesh snapshot
plus every expected output the reference produced into one compressed .npz per case.
Must be run in the build container (needs /root/reference for the PNG stack and the binary);
the fixtures it writes are committed so that the GPU box, which has neither, can still check
the oracle and the CUDA path against reference-produced numbers.

usage: python tests/golden/make_golden.py [case ...]
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dsc_ref")
REF_DATA = "/root/reference/data"


def png_stack(name):
    from PIL import Image
    d = os.path.join(REF_DATA, name)
    n = len([f for f in os.listdir(d) if f.lower().endswith(".png")])
    return np.stack([np.array(Image.open(os.path.join(d, "im_%d.png" % i))) for i in range(n)]).astype(np.uint8)


def sphere_f32(n, radius, lo, hi):
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    c = (n - 1) / 2.0
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    return np.where(r < radius, hi, lo).astype(np.float32)


def shells_u16(n, seed, sigma):
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    c = (n - 1) / 2.0
    r = np.sqrt((x - c) ** 2 + (y - c) ** 2 + (z - c) ** 2)
    v = np.where(r < 0.22 * n, 0.9, np.where(r < 0.40 * n, 0.5, 0.1))
    rng = np.random.default_rng(seed)
    v = np.clip(v + rng.normal(0, sigma, v.shape), 0, 1)
    return np.round(v * 65535).astype(np.uint16)


DT = {"u8": np.uint8, "u16": np.uint16, "f32": np.float32, "f64": np.float64}
EXT2DT = {"f64": np.float64, "u8": np.uint8, "u32": np.uint32, "i32": np.int32, "u64": np.uint64}

CASES = {
    # BASELINE config #1: square_round_2.txt (edge 6, 3 phases, alpha 0.001), Otsu labels
    "c1_e6": dict(vol=lambda: png_stack("square_round_2"), dtype="u8", edge=6, nb_phase=3, alpha=0.001, labels="otsu"),
    # same volume after two full segment() iterations (deform + relabel + adapt at iter 0): an irregular, deformed mesh
    "c1_e8_it2": dict(vol=lambda: png_stack("square_round_2"), dtype="u8", edge=8, nb_phase=3, alpha=0.001, labels="otsu", iters=2),
    # noisy variant of the reference's data, jittered mesh, fixed thresholds
    "c1n_e9_j": dict(vol=lambda: png_stack("square_round_2_0.1"), dtype="u8", edge=9, nb_phase=3, alpha=0.01, labels="thres",
                     thres="0.1,0.45", jitter=0.22),
    # float sphere (BASELINE config #2 scaled down), 2 phases
    "sph48_f32": dict(vol=lambda: sphere_f32(48, 15.0, 0.2, 0.8), dtype="f32", edge=4.3, nb_phase=2, alpha=0.001, labels="thres",
                      thres="0.5", jitter=0.25),
    # non-cubic volume (X=100, Y=72, Z=56: a crop of the reference's data), 4 phases by fixed thresholds, jittered mesh
    "c1_crop_xyz": dict(vol=lambda: np.ascontiguousarray(png_stack("square_round_2_0.1")[22:78, 14:86, :]), dtype="u8", edge=7.3, nb_phase=4,
                        alpha=0.005, labels="thres", thres="0.08,0.3,0.6", jitter=0.2, iters=1),
    # uint16 nested shells + noise (BASELINE config #3 scaled down), coarse mesh => higher sample levels
    "sh40_u16": dict(vol=lambda: shells_u16(40, 1234, 0.05), dtype="u16", edge=10, nb_phase=3, alpha=0.001, labels="thres",
                     thres="0.3,0.7", jitter=0.2),
}


def run_case(name, spec):
    vol = spec["vol"]()
    assert vol.dtype == DT[spec["dtype"]]
    with tempfile.TemporaryDirectory() as tmp:
        vpath = os.path.join(tmp, "vol.raw")
        vol.tofile(vpath)
        out = os.path.join(tmp, "out")
        cmd = [REF_BIN, "--mode", "dump", "--volume", vpath, "--vol-dtype", spec["dtype"], "--dims", str(vol.shape[2]), str(vol.shape[1]),
               str(vol.shape[0]), "--edge", str(spec["edge"]), "--nb-phase", str(spec["nb_phase"]), "--alpha", str(spec["alpha"]), "--labels",
               spec["labels"], "--out", out, "--points-in-tet", "1500", "--iters", str(spec.get("iters", 0)), "--jitter",
               str(spec.get("jitter", 0))]
        if "thres" in spec:
            cmd += ["--thres", spec["thres"]]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
        meta = json.load(open(os.path.join(out, "meta.json")))
        arrs = {"volume": vol}
        for f in os.listdir(out):
            if f == "meta.json":
                continue
            base, ext = f.rsplit(".", 1)
            arrs[base] = np.fromfile(os.path.join(out, f), dtype=EXT2DT[ext])
    P = meta["nb_phase"]
    arrs["pos"] = arrs["pos"].reshape(-1, 3)
    arrs["forces"] = arrs["forces"].reshape(-1, 3)
    arrs["internal_forces"] = arrs["internal_forces"].reshape(-1, 3)
    arrs["tet_nodes"] = arrs["tet_nodes"].reshape(-1, 4)
    arrs["face_nodes"] = arrs["face_nodes"].reshape(-1, 3)
    arrs["face_tets"] = arrs["face_tets"].reshape(-1, 2)
    arrs["pit_pts"] = arrs["pit_pts"].reshape(-1, 3)
    arrs["tet_face_nodes"] = arrs["tet_face_nodes"].reshape(-1, 4, 3)
    arrs["tet_face_tets"] = arrs["tet_face_tets"].reshape(-1, 4, 2)
    for k in ("tet_variation", "tet_energy", "tet_relabel_energy"):
        arrs[k] = arrs[k].reshape(-1, P)
    arrs["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-12s %7d tets %6d interface faces  mean=%s  -> %s (%.0f KB)" % (
        name, meta["n_tets"], meta["n_if"], np.array2string(arrs["mean"], precision=10), os.path.relpath(path, ROOT),
        os.path.getsize(path) / 1024))


def main():
    names = sys.argv[1:] or list(CASES)
    if not os.path.exists(REF_BIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    for n in names:
        run_case(n, CASES[n])


if __name__ == "__main__":
    main()
