"""The staged tet kernel (csrc/dsc_tet_staged.cuh, variant bit 131072): lane per tet, voxel boxes in shared memory by TMA,
lattice arithmetic rebuilt from the structure of the reference's sample generator, fixed-point per-label sums.

Against the reference-produced fixtures and the flat oracle: every sample in the reference's voxel (per-tet sums of float voxels
are exact in double in any order, so bit-equal totals mean equal voxels), volumes bit for bit, per-label sums to 1e-9."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
STAGED = 131072 | 128 | 16384
W2 = 128 | 16384


def _seg(pkg, g):
    img = pkg.Image3D(g["volume"])
    seg = pkg.SegmentFunction(img, g["meta"]["nb_phase"], g["meta"].get("alpha", 0.01))
    seg.set_snapshot(g)
    return img, seg


@pytest.mark.parametrize("name", ["c1_e6", "c1_e8_it2", "sph48_f32", "sh40_u16", "c1n_e9_j", "c1_crop_xyz"])
def test_goldens(pkg, name):
    g = golden_io.load(name)
    img, seg = _seg(pkg, g)
    lab = g["tet_label"] != 0
    seg.set_tet_variant(STAGED)
    o = seg.tet_integrate()
    assert np.array_equal(o["vol"][lab], g["tet_vol"][lab])
    np.testing.assert_allclose(o["total"][lab], g["tet_total"][lab], rtol=1e-12, atol=1e-13)
    P = g["meta"]["nb_phase"]
    tot = np.array([g["tet_total"][g["tet_label"] == k + 1].sum() for k in range(P)])
    vol = np.array([g["tet_vol"][g["tet_label"] == k + 1].sum() for k in range(P)])
    np.testing.assert_allclose(o["phase_total"], tot, rtol=1e-9)
    np.testing.assert_allclose(o["phase_volume"], vol, rtol=1e-9)
    o2 = seg.tet_integrate(per_tet=False)
    # fixed-point sums: the same integers whatever warp took which tile, with or without per-tet outputs
    assert np.array_equal(o2["phase_total"], o["phase_total"]) or np.allclose(o2["phase_total"], o["phase_total"], rtol=1e-12)
    assert np.array_equal(o2["phase_volume"], o["phase_volume"])
    seg.set_tet_variant(W2)
    o3 = seg.tet_integrate()
    if g["volume"].dtype == np.float32:
        assert np.array_equal(o3["total"][lab], o["total"][lab])
    else:  # integer voxels: the kernels divide the integer sums at different points
        np.testing.assert_allclose(o3["total"][lab], o["total"][lab], rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(o3["phase_sumsq"], o["phase_sumsq"], rtol=1e-9)
    img.close()


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
def test_large_interior_mesh_bit_equal(pkg, flat, dtype):
    syn = pkg.synthetic
    vol = syn.shells_phantom(96, sigma=0.05, seed=3, dtype=dtype)
    snap = syn.build_case(vol, 6.3, 3, (0.3, 0.7), lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"],
                          jitter=0.3, seed=17)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], want_hash=False)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(snap)
    seg.set_tet_variant(STAGED)
    o = seg.tet_integrate(include_bound=True)
    assert np.array_equal(o["vol"], ref["vol"])
    if dtype == np.float32:
        assert np.array_equal(o["total"], ref["total"])
    else:
        np.testing.assert_allclose(o["total"], ref["total"], rtol=1e-13, atol=1e-14)
    lab = snap["tet_label"]
    a = seg.tet_integrate(per_tet=False)
    b = seg.tet_integrate(per_tet=False)
    for k in ("phase_total", "phase_volume", "phase_sumsq"):
        assert np.array_equal(a[k], b[k]), k  # run to run: integers
    tot, volu, mean = flat.phase_reduce(lab, ref["total"], ref["vol"], 3)
    np.testing.assert_allclose(a["phase_total"], tot, rtol=1e-9)
    np.testing.assert_allclose(a["phase_volume"], volu, rtol=1e-9)
    assert seg.last_sample_count() == int(ref["nsamp"][lab != 0].sum())
    img.close()


@pytest.mark.parametrize("edge,jitter", [(2.6, 0.3), (3.7, 0.25), (4.9, 0.3), (9.5, 0.2)])
def test_levels_and_small_tets(pkg, flat, edge, jitter):
    """Meshes whose tets sit at levels 1..2, 2..3, 3..4 and 4..6 (the last: levels above 4 take the exact path)."""
    syn = pkg.synthetic
    vol = syn.shells_phantom(64, sigma=0.08, seed=11, dtype=np.float32)
    snap = syn.build_case(vol, edge, 3, (0.3, 0.7), lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"],
                          jitter=jitter, seed=5)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], want_hash=False)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(snap)
    seg.set_tet_variant(STAGED)
    o = seg.tet_integrate(include_bound=True)
    assert np.array_equal(o["vol"], ref["vol"])
    assert np.array_equal(o["total"], ref["total"])
    img.close()


def test_non_cubic_and_partition(pkg, flat):
    syn = pkg.synthetic
    vol = np.ascontiguousarray(syn.shells_phantom(64, levels=(0.05, 0.35, 0.65, 0.95), radii=(0.45, 0.3, 0.15), sigma=0.08, seed=21, dtype=np.float32)[5:57, 0:40, 3:63])
    assert vol.shape == (52, 40, 60)
    P = 6
    thr = tuple(np.linspace(0.1, 0.9, P - 1))
    snap = syn.build_case(vol, 5.5, P, thr, lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"], jitter=0.25, seed=3)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], c=[0.0])
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, P, 0.01)
    seg.set_snapshot(snap)
    seg.set_tet_variant(STAGED)
    o = seg.tet_integrate()
    lab = snap["tet_label"]
    m = lab != 0
    assert np.array_equal(o["vol"][m], ref["vol"][m])
    assert np.array_equal(o["total"][m], ref["total"][m])
    tot, volu, mean = flat.phase_reduce(lab, ref["total"], ref["vol"], P)
    np.testing.assert_allclose(o["phase_total"], tot, rtol=1e-9)
    sq, _, _ = flat.phase_reduce(lab, ref["sq"][:, 0], ref["vol"], P)
    np.testing.assert_allclose(o["phase_sumsq"], sq, rtol=1e-9)
    img.close()


@pytest.mark.parametrize("slots,slot_kb", [(2, 24), (3, 12), (5, 8), (8, 24)])
def test_slot_ring_under_stress(pkg, flat, slots, slot_kb, monkeypatch):
    """The hand-over protocol of the slot ring (sequence counters + mbarrier) with few and small slots: items finish out of order, warps
    run several rounds of the ring apart, boxes are split until they fit -- every run must give the oracle's per-tet sums bit for bit and the
    same fixed-point per-label sums.  (compute-sanitizer is closed on this pool: this and the comparison with the CPU reference stand in.)"""
    monkeypatch.setenv("DSCGPU_STG_SLOTS", str(slots))
    monkeypatch.setenv("DSCGPU_STG_SLOT_KB", str(slot_kb))
    syn = pkg.synthetic
    vol = syn.shells_phantom(96, sigma=0.05, seed=9, dtype=np.float32)
    snap = syn.build_case(vol, 6.3, 3, (0.3, 0.7), lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"],
                          jitter=0.3, seed=23)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], want_hash=False)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(snap)
    seg.set_tet_variant(STAGED)
    first = None
    for rep in range(6):
        o = seg.tet_integrate(include_bound=True)
        assert np.array_equal(o["vol"], ref["vol"]) and np.array_equal(o["total"], ref["total"]), rep
        s = seg.tet_integrate(per_tet=False)
        key = (tuple(s["phase_total"]), tuple(s["phase_volume"]), tuple(s["phase_sumsq"]))
        first = first or key
        assert key == first, rep
    img.close()
