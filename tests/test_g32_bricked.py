"""The warp-per-tet kernel over the bricked volume (csrc/dsc_tet_g32.cuh, tet variant bit 1024 -- the default of the hot
path) against the flat oracle.

Float voxel sums of <= 729 terms are exact in double whatever the order, so BIT-EQUAL per-tet totals mean that every
sample landed in the same voxel as in the reference (image3d::get_tetra_intensity, DEMO/image3d.cpp:351-382).  The cases
aim at what is new in this kernel: brick addressing with brick-aligned origins, the zero / replicated padding that stands
in for image3d::get_value's bounds test and the truncation toward zero of coordinates in (-1,0), tets of every sample
level sharing a tile, label runs, the slow path, and volumes that are not "simple" (zeros, negatives, denormals).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-6
G32 = 1024 | 128
W2B = 2048 | 128  # w2 gathering from the bricked copy
W2P = 8192 | 128  # w2 with packed FP32 position arithmetic (hot-path instantiation only)
W2PF = 65536 | 128  # w2 with the TMA L2 prefetch of the next tile's voxel box (hot-path instantiation only; a hint: same results)


def _random_tets(rng, dims_xyz, n_tets, size_lo, size_hi, margin):
    """Independent random tetrahedra: a centre inside [-margin, dim+margin) and four corners within +-size of it."""
    X, Y, Z = dims_xyz
    c = rng.uniform([-margin] * 3, [X + margin, Y + margin, Z + margin], (n_tets, 1, 3))
    size = rng.uniform(size_lo, size_hi, (n_tets, 1, 1))
    pos = (c + rng.uniform(-1.0, 1.0, (n_tets, 4, 3)) * size).reshape(-1, 3)
    tets = np.arange(4 * n_tets, dtype=np.uint32).reshape(-1, 4)
    return pos, tets


def _snap(pos, tets, labels):
    return {"pos": pos, "tet_nodes": tets, "tet_label": labels.astype(np.int32), "face_nodes": np.zeros((0, 3), np.uint32),
            "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)}


def _check(pkg, flat, vol, pos, tets, labels, P, exact=True):
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, P)
    seg.set_snapshot(_snap(pos, tets, labels))
    ref = flat.tet_integrate(vol, tets, pos, c=[0.0], want_hash=False)
    m = labels != 0
    outs = {}
    for variant in (G32, W2B, 128):
        seg.set_tet_variant(variant)
        o = seg.tet_integrate()
        assert np.array_equal(o["vol"][m], ref["vol"][m])
        if exact:
            assert np.array_equal(o["total"][m], ref["total"][m]), variant
        else:
            np.testing.assert_allclose(o["total"][m], ref["total"][m], rtol=1e-12, atol=1e-13)
        assert np.all(o["total"][~m] == -1.0)
        tot, volu, _ = flat.phase_reduce(labels, ref["total"], ref["vol"], P)
        sq, _, _ = flat.phase_reduce(labels, ref["sq"][:, 0], ref["vol"], P)
        scale = np.abs(tot).max() + 1e-300
        np.testing.assert_allclose(o["phase_total"], tot, rtol=RTOL, atol=1e-12 * scale)
        np.testing.assert_allclose(o["phase_volume"], volu, rtol=RTOL)
        np.testing.assert_allclose(o["phase_sumsq"], sq, rtol=RTOL, atol=1e-12 * scale)
        o2 = seg.tet_integrate(per_tet=False)  # the instantiation the hot path runs: no per-tet outputs
        np.testing.assert_allclose(o2["phase_total"], tot, rtol=RTOL, atol=1e-12 * scale)
        np.testing.assert_allclose(o2["phase_volume"], volu, rtol=RTOL)
        outs[variant] = o
    # the hot path proper (dscgpu_tet_phase_sums_dev: no second moment, no per-tet outputs) through the fused evaluation
    for variant in (G32, W2B, W2P, W2PF, 128 | 4096, 128 | 16384, 128 | 16384 | 32768, 128 | 8192 | 16384):
        seg.set_tet_variant(variant)
        mean, _ = seg.image_force_eval(_snap(pos, tets, labels), mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_ATOMIC)
        ok = volu > 0
        np.testing.assert_allclose(np.asarray(mean)[ok], (tot / np.where(ok, volu, 1))[ok], rtol=RTOL, atol=1e-12)
        assert seg.last_sample_count() == int(ref["nsamp"][m].sum())
    img.close()
    return outs


@pytest.mark.parametrize("seed,dims,size", [(1, (40, 36, 44), (1.0, 9.0)), (2, (33, 57, 20), (0.5, 4.0)), (3, (64, 64, 64), (4.0, 16.0))])
def test_random_tets_all_levels_and_borders(pkg, flat, seed, dims, size):
    """Random tets of every sample level (n = 1..9) anywhere from well outside the volume to its interior; labels in
    random order (0 = skipped, 1..P, and P+1 which is processed but belongs to no phase)."""
    rng = np.random.default_rng(seed)
    X, Y, Z = dims
    vol = rng.uniform(0.05, 1.0, (Z, Y, X)).astype(np.float32)
    pos, tets = _random_tets(rng, dims, 6000, size[0], size[1], margin=12.0)
    P = 4
    labels = rng.integers(0, P + 2, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, P)


def test_levels_are_all_exercised(pkg, flat):
    rng = np.random.default_rng(5)
    vol = rng.uniform(0.05, 1.0, (48, 48, 48)).astype(np.float32)
    pos, tets = _random_tets(rng, (48, 48, 48), 4000, 1.0, 14.0, margin=2.0)
    ref = flat.tet_integrate(vol, tets, pos, want_hash=False)
    levels = np.unique(np.rint(np.cbrt(ref["nsamp"])).astype(int))
    assert set(levels) == set(range(1, 10))
    labels = rng.integers(1, 4, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, 3)


def test_coordinates_between_minus_one_and_zero_read_voxel_zero(pkg, flat):
    """(int)(-0.4) == 0: samples just outside the low faces read layer 0 (replicated padding), samples below -1 and at or
    beyond dim read 0; thin tets hugging every face of the volume."""
    rng = np.random.default_rng(11)
    X, Y, Z = 24, 20, 28
    vol = rng.uniform(0.1, 1.0, (Z, Y, X)).astype(np.float32)
    pos_l, tets_l = [], []
    for axis in range(3):
        for side, dim in ((0, None), (1, (X, Y, Z)[axis])):
            p, t = _random_tets(rng, (X, Y, Z), 800, 0.8, 5.0, margin=0.0)
            lo = -1.6 if side == 0 else dim - 1.4
            p[:, axis] = lo + (p[:, axis] % 3.0)  # squeeze the tets into a 3-voxel slab straddling the face
            tets_l.append(t + 4 * 800 * len(pos_l))
            pos_l.append(p)
    pos, tets = np.concatenate(pos_l), np.concatenate(tets_l).astype(np.uint32)
    labels = rng.integers(1, 3, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, 2)


def test_slow_path_wide_and_far_tets(pkg, flat):
    """Tets wider than 100 voxels or outside the padded copy take the exact whole-warp path."""
    rng = np.random.default_rng(13)
    vol = rng.uniform(0.1, 1.0, (130, 40, 150)).astype(np.float32)
    pos, tets = _random_tets(rng, (150, 40, 130), 600, 3.0, 8.0, margin=30.0)
    # a few needle tets spanning > 100 voxels with a small volume
    for k in range(0, 200, 4):
        pos[4 * k] = [1.0 + k * 0.1, 5.0, 2.0]
        pos[4 * k + 1] = [140.0, 6.0 + k * 0.01, 3.0]
        pos[4 * k + 2] = [70.0, 7.5, 4.0 + k * 0.02]
        pos[4 * k + 3] = [71.0, 5.5, 120.0]
    labels = rng.integers(1, 4, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, 3)


def test_non_simple_float_volume(pkg, flat):
    """Zeros, negative values and denormals in a float volume: the general widening routine, same results."""
    rng = np.random.default_rng(17)
    vol = rng.uniform(-0.5, 1.0, (32, 32, 32)).astype(np.float32)
    vol[rng.uniform(size=vol.shape) < 0.3] = 0.0
    vol[3, 4, 5] = np.float32(1e-41)  # denormal
    pos, tets = _random_tets(rng, (32, 32, 32), 3000, 1.0, 7.0, margin=3.0)
    labels = rng.integers(0, 4, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, 3)
    vol0 = np.zeros((32, 32, 32), np.float32)  # all zero: sums must be exactly 0, not denormal residue
    outs = _check(pkg, flat, vol0, pos, tets, labels, 3)
    assert np.all(outs[G32]["phase_total"] == 0.0)
    volz = np.abs(vol)  # "simple" but with exact zeros
    _check(pkg, flat, volz, pos, tets, labels, 3)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float64])
def test_other_voxel_types(pkg, flat, dtype):
    rng = np.random.default_rng(19)
    if dtype == np.float64:
        vol = rng.uniform(0.0, 1.0, (30, 34, 38))
    else:
        vol = rng.integers(0, np.iinfo(dtype).max + 1, (30, 34, 38)).astype(dtype)
    pos, tets = _random_tets(rng, (38, 34, 30), 4000, 1.0, 9.0, margin=6.0)
    labels = rng.integers(0, 9, len(tets))
    _check(pkg, flat, vol, pos, tets, labels, 8, exact=False)


def test_grid_mesh_interior_matches_w2_bitwise(pkg, flat):
    """A jittered grid mesh like the benchmark's (neighbouring tets share tiles, most tets interior)."""
    syn = pkg.synthetic
    vol = syn.shells_phantom(96, sigma=0.05, seed=3, dtype=np.float32)
    snap = syn.build_case(vol, 6.3, 3, (0.3, 0.7), lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"],
                          jitter=0.3, seed=17)
    _check(pkg, flat, vol, snap["pos"], snap["tet_nodes"], snap["tet_label"], 3)


def test_partition_ranges_compose(pkg, flat):
    """Multi-GPU contract on the new kernel: ranges that do not start on a tile boundary."""
    rng = np.random.default_rng(23)
    vol = rng.uniform(0.05, 1.0, (40, 40, 40)).astype(np.float32)
    pos, tets = _random_tets(rng, (40, 40, 40), 5000, 1.0, 8.0, margin=4.0)
    labels = rng.integers(0, 4, len(tets))
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(_snap(pos, tets, labels))
    seg.set_tet_variant(G32)
    whole = seg.tet_integrate(per_tet=False)
    tot = np.zeros(3)
    volu = np.zeros(3)
    cuts = [0, 1, 33, 1777, 1778, 4096, 5000]
    for b, e in zip(cuts[:-1], cuts[1:]):
        seg.set_partition((b, e), (0, 0), (0, seg.n_nodes_buffer))
        o = seg.tet_integrate(per_tet=False)
        tot += o["phase_total"]
        volu += o["phase_volume"]
    seg.set_partition((0, len(tets)), (0, 0), (0, seg.n_nodes_buffer))
    np.testing.assert_allclose(tot, whole["phase_total"], rtol=1e-12)
    np.testing.assert_allclose(volu, whole["phase_volume"], rtol=1e-12)
    img.close()


def test_volume_modified_rebuilds_the_bricked_copy(pkg, flat):
    rng = np.random.default_rng(29)
    vol = rng.uniform(0.05, 1.0, (24, 24, 24)).astype(np.float32)
    pos, tets = _random_tets(rng, (24, 24, 24), 1000, 1.0, 6.0, margin=2.0)
    labels = rng.integers(1, 3, len(tets))
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 2)
    seg.set_snapshot(_snap(pos, tets, labels))
    seg.set_tet_variant(G32)
    a = seg.tet_integrate()
    vol2 = (vol * np.float32(0.5) + np.float32(0.25)).astype(np.float32)
    img.write_volume(vol2)
    b = seg.tet_integrate()
    ref = flat.tet_integrate(vol2, tets, pos, want_hash=False)
    assert np.array_equal(b["total"], ref["total"])
    assert not np.array_equal(a["total"], b["total"])
    img.close()


def test_l2_prefetch_variant_on_many_tiles(pkg, flat):
    """Variant bit 65536 only adds TMA L2 prefetches of the next tile's voxel box (a hint).  It engages when a launch has
    more 32-tet tiles than resident warps, i.e. above ~114 K tets: the sums must be bit-identical to plain w2's."""
    syn = pkg.synthetic
    vol = syn.shells_phantom(160, sigma=0.05, seed=5, dtype=np.float32)
    pos, tets, n_grid = syn.grid_mesh(vol.shape[::-1], 5.9, jitter=0.25, seed=9)
    assert len(tets) > 130000
    labels = syn.boundary_layer_labels(tets, n_grid) * (1 + (np.arange(len(tets)) % 3))
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(_snap(pos, tets, labels))
    outs = {}
    for variant in (128, W2PF):
        seg.set_tet_variant(variant)
        for _ in range(2):  # second pass: tile boxes cached
            mean, _ = seg.image_force_eval(None, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_ATOMIC)
        outs[variant] = (np.array(mean), seg.last_sample_count())
    assert np.array_equal(outs[128][0], outs[W2PF][0]) and outs[128][1] == outs[W2PF][1]
    ref = flat.tet_integrate(vol, tets, pos, want_hash=False)
    tot, volu, _ = flat.phase_reduce(labels, ref["total"], ref["vol"], 3)
    np.testing.assert_allclose(outs[W2PF][0], tot / volu, rtol=RTOL)
    img.close()


def test_nodes_exactly_on_the_volume_faces(pkg, flat):
    """The first layer of tets of a DSC mesh has nodes exactly on the faces of the volume (coordinate 0.0 or dim).  w2 takes
    such tets on its FP32-filtered path when no sample can leave the volume; slivers along a face stay on the exact path.
    Every sample must still land in the reference's voxel (bit-equal per-tet totals), whichever path a tet took."""
    rng = np.random.default_rng(77)
    dims = (48, 40, 56)
    vol = rng.uniform(0.05, 1.0, dims[::-1]).astype(np.float32)
    pos, tets = _random_tets(rng, dims, 6000, 2.0, 9.0, margin=0.0)
    pos = np.clip(pos, 0.0, np.array(dims, dtype=np.float64))  # many corners land exactly on a face
    # slivers: all four corners within 1e-3 (or exactly on) of the top face of one axis
    for k in range(600):
        ax = k % 3
        t = tets[k]
        pos[t, ax] = dims[ax] - rng.uniform(0.0, 1.0e-3, 4) * (k % 2)
    # and a negative-zero / tiny negative corner, which truncates to voxel 0
    pos[tets[600:700, 0], 0] = -1.0e-15
    labels = rng.integers(1, 4, len(tets))
    outs = _check(pkg, flat, vol, pos, tets, labels, 3)
    assert np.array_equal(outs[128]["total"], outs[W2B]["total"])
