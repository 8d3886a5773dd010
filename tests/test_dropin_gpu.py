"""Drop-in on the GPU box: the unmodified reference objects and the product's C ABI in one process, on a live,
deforming is_mesh/DSC complex (oracle/dropin_test.cpp), plus a fresh reference run (not a committed fixture)
compared through the ctypes path."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dsc_dropin")
REF = os.path.join(ROOT, "oracle", "_ref", "dsc_ref")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/dsc_dropin is built in the container (make -C oracle dropin)")
def test_reference_segment_loop_with_gpu_dropin():
    g = golden_io.load("c1_e6")
    with tempfile.TemporaryDirectory() as tmp:
        vp = os.path.join(tmp, "vol.u8")
        g["volume"].tofile(vp)
        # edge 8, two full segment() iterations (deform + relabel + adapt) between the comparisons
        r = subprocess.run([DROPIN, vp, "100", "100", "100", "8", "2", "3", "0.001"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DROP-IN PARITY OK" in r.stdout
    assert r.stdout.count("phase_volume bit-exact") == 3
    assert r.stdout.count("destinations bit-exact") == 3  # get_node_displacement + snapp_boundary on the reference's forces
    assert "0 labels differ" in r.stdout                  # initialisation pass: no tolerance


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/dsc_dropin is built in the container (make -C oracle dropin)")
def test_dropin_on_a_quarter_million_tets_with_timings():
    """SURVEY section 6, edge-3 row: 257 250 tets / 23.5 K interface faces on a live is_mesh complex; parity as above, and the
    wall-clock of every reference body next to its drop-in (the is_mesh walk of ImageEnergy::refresh included) in the log."""
    g = golden_io.load("c1_e6")
    with tempfile.TemporaryDirectory() as tmp:
        vp = os.path.join(tmp, "vol.u8")
        g["volume"].tofile(vp)
        r = subprocess.run([DROPIN, vp, "100", "100", "100", "3", "1", "3", "0.001"], capture_output=True, text=True, timeout=1500)
    print(r.stdout[-4000:], r.stderr[-2000:])
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "dropin_edge3.log"), "w") as f:
            f.write(r.stdout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DROP-IN PARITY OK" in r.stdout
    assert "timing [ms] reference" in r.stdout


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/dsc_ref is built in the container (make -C oracle ref)")
def test_fresh_reference_run_matches_gpu(pkg):
    rng_seed = 4242
    vol = pkg.synthetic.shells_phantom(44, sigma=0.06, seed=rng_seed, dtype=np.uint16)
    with tempfile.TemporaryDirectory() as tmp:
        vp = os.path.join(tmp, "vol.raw")
        vol.tofile(vp)
        out = os.path.join(tmp, "out")
        subprocess.run([REF, "--mode", "dump", "--volume", vp, "--vol-dtype", "u16", "--dims", "44", "44", "44", "--edge", "6.5", "--nb-phase", "3",
                        "--alpha", "0.02", "--labels", "thres", "--thres", "0.3,0.7", "--jitter", "0.28", "--jitter-seed", str(rng_seed), "--out", out,
                        "--iters", "1"], check=True, capture_output=True, timeout=600)
        meta = json.load(open(os.path.join(out, "meta.json")))
        rd = lambda n, dt: np.fromfile(os.path.join(out, n), dtype=dt)
        snap = {"pos": rd("pos.f64", np.float64).reshape(-1, 3), "tet_nodes": rd("tet_nodes.u32", np.uint32).reshape(-1, 4),
                "tet_label": rd("tet_label.i32", np.int32), "face_nodes": rd("face_nodes.u32", np.uint32).reshape(-1, 3),
                "face_tets": rd("face_tets.u32", np.uint32).reshape(-1, 2), "face_is_boundary": rd("face_is_boundary.u8", np.uint8)}
        ref = {k: rd(k + ".f64", np.float64) for k in ("mean", "total", "phase_volume", "forces", "tet_total", "tet_vol")}
        ref_hash = rd("tet_hash.u64", np.uint64)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3, 0.02)
    seg.set_snapshot(snap)
    h, _ = seg.tet_sample_hash()
    assert np.array_equal(h, ref_hash)
    mean = seg.update_average_intensity()
    assert np.array_equal(seg._phase_volume, ref["phase_volume"])
    np.testing.assert_allclose(mean, ref["mean"], rtol=1e-6)
    F = seg.compute_external_force(mean=ref["mean"])
    assert np.array_equal(F, ref["forces"].reshape(-1, 3))
    d, a = seg.compute_energy(mean=ref["mean"])
    np.testing.assert_allclose([d, a], [meta["energy_data"], meta["energy_area"]], rtol=1e-6)
    lab = snap["tet_label"] != 0
    o = seg.tet_integrate()
    assert np.array_equal(o["vol"][lab], ref["tet_vol"][lab])
    np.testing.assert_allclose(o["total"][lab], ref["tet_total"][lab], rtol=1e-6)
    img.close()
