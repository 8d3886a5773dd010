"""Parity of the CUDA path (through the C ABI) with the reference: against the reference-produced
golden fixtures and, on seeded synthetic inputs, against the flat oracle.

Bars (BASELINE.json north_star): bit-exact for sample->voxel classification (hash), sample levels,
voxel counts (_phase_volume), the sum table, vertex_bound, the voxel-in-tet predicate and -- because
the gather kernel keeps the reference's addition order -- the forces; 1e-6 relative (FP64 parallel
sums) for totals, means and energies; 1e-6 for the atomic force kernel.
"""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
RTOL = 1e-6  # FP64 accumulation tolerance stated by north_star


def _seg(pkg, g):
    img = pkg.Image3D(g["volume"])
    seg = pkg.SegmentFunction(img, g["meta"]["nb_phase"], g["meta"]["alpha"])
    seg.set_snapshot(golden_io.snapshot_of(g))
    return img, seg


@pytest.fixture(scope="module", params=golden_io.CASES)
def case(request, pkg):
    g = golden_io.load(request.param)
    img, seg = _seg(pkg, g)
    yield g, img, seg
    img.close()


def test_sample_classification_bit_exact(case):
    g, img, seg = case
    h, n = seg.tet_sample_hash()
    assert np.array_equal(n, g["tet_nsamp"])
    assert np.array_equal(h, g["tet_hash"])


@pytest.mark.parametrize("lanes", [8, 32])
def test_tet_integrate_matches_reference(case, lanes):
    g, img, seg = case
    seg.set_lanes_per_tet(lanes)
    lab = g["tet_label"] != 0
    o = seg.tet_integrate(c=g["mean"])
    assert np.array_equal(o["vol"][lab], g["tet_vol"][lab])  # volumes are computed in the reference's order: bit-exact
    np.testing.assert_allclose(o["total"][lab], g["tet_total"][lab], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o["mean"][lab], g["tet_mean"][lab], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o["variation"][lab], g["tet_variation"][lab], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o["energy"][lab], g["tet_energy"][lab], rtol=RTOL, atol=1e-12)
    assert np.all(o["total"][~lab] == -1.0)
    # per-label reduction vs the sequential reduction of the reference's per-tet values
    P = g["meta"]["nb_phase"]
    tot = np.array([g["tet_total"][g["tet_label"] == k + 1].sum() for k in range(P)])
    vol = np.array([g["tet_vol"][g["tet_label"] == k + 1].sum() for k in range(P)])
    np.testing.assert_allclose(o["phase_total"], tot, rtol=RTOL)
    np.testing.assert_allclose(o["phase_volume"], vol, rtol=RTOL)
    np.testing.assert_allclose(o["phase_mean"], tot / vol, rtol=RTOL)
    seg.set_lanes_per_tet(0)


def test_tet_integrate_fast_path_equals_full_path(case):
    g, img, seg = case
    a = seg.tet_integrate(c=None)
    b = seg.tet_integrate(c=g["mean"])
    np.testing.assert_allclose(a["total"], b["total"], rtol=1e-12)
    np.testing.assert_allclose(a["phase_sumsq"], b["phase_sumsq"], rtol=1e-9)


def test_sum_table_bit_exact(case, flat):
    g, img, seg = case
    img.build_sum_table()
    assert np.array_equal(img.sum_table(), flat.build_sum_table(g["volume"]))


def test_update_average_intensity(case):
    g, img, seg = case
    mean = seg.update_average_intensity()
    assert np.array_equal(seg._phase_volume, g["phase_volume"])  # voxel counts: bit-exact
    np.testing.assert_allclose(seg._total_intensities, g["total"], rtol=RTOL)
    np.testing.assert_allclose(mean, g["mean"], rtol=RTOL)


def test_compute_external_force_gather_bit_exact(case, pkg):
    g, img, seg = case
    F = seg.compute_external_force(mode=pkg.FORCE_GATHER, mean=g["mean"])
    assert np.array_equal(F, g["forces"])
    # the mean-independent half started ahead on the second stream (as next to the tet pass), joined by the force call
    seg.face_geometry_async()
    seg.tet_integrate(per_tet=False)
    F2 = seg.compute_external_force(mode=pkg.FORCE_GATHER, mean=g["mean"])
    assert np.array_equal(F2, g["forces"])


def test_compute_external_force_atomic(case, pkg):
    g, img, seg = case
    F = seg.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=g["mean"])
    scale = np.abs(g["forces"]).max()
    np.testing.assert_allclose(F, g["forces"], rtol=RTOL, atol=RTOL * scale)


def test_compute_energy(case):
    g, img, seg = case
    vb = seg.vertex_bound()
    assert np.array_equal(vb, g["vertex_bound"])
    d, a = seg.compute_energy(mean=g["mean"])
    np.testing.assert_allclose(d, g["meta"]["energy_data"], rtol=RTOL)
    np.testing.assert_allclose(a, g["meta"]["energy_area"], rtol=RTOL)
    d2, a2 = seg.compute_energy(mean=g["mean"], vertex_bound=g["vertex_bound"])
    assert d2 == d and a2 == a


def test_initialisation_pass(case, flat):
    """histogram of the per-tet means and the label rule against the reference-derived fixture values."""
    g, img, seg = case
    h = seg.init_histogram()
    ref = g["init_hist"]
    assert h.sum() == ref.sum() == int((g["tet_label"] != 0).sum())
    # the per-tet means of this pass are added in the reference's order (tet_mean_ordered_kernel): bins and labels are truncations of
    # bit-identical numbers -- no tolerance, also for tets whose mean sits on a bin edge
    assert np.array_equal(h, ref)
    lab = seg.init_labels(g["init_thres"])
    assert np.array_equal(lab, g["init_labels"])
    assert np.all(lab[g["tet_label"] == 0] == 0)
    assert np.array_equal(seg.init_labels([]), np.where(g["tet_label"] != 0, 1, 0))  # one phase: no threshold


def test_internal_force(case, flat):
    """compute_internal_force_2 against the reference's own _internal_forces."""
    g, img, seg = case
    fl = g["node_flags"]
    ref = g["internal_forces"]
    F = seg.compute_internal_force_2(fl)  # m_vertex_bound derived on the device
    unprojected = ((fl & 1) == 0) | ((fl & 2) != 0)
    assert np.array_equal(F[unprojected], ref[unprojected])  # ordered sums, reference arithmetic: bit for bit
    np.testing.assert_allclose(F, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    F2 = seg.compute_internal_force_2(fl, vertex_bound=g["vertex_bound"])
    assert np.array_equal(F2, F)
    assert np.array_equal(F, flat.internal_forces(golden_io.snapshot_of(g), fl, g["vertex_bound"]))  # same order as the flat oracle
    # with the get_faces(node) order of the reference's complex (DSC::get_normal, src/DSC.h:3138-3159): every node bit for bit
    order = (g["node_face_off"], g["node_face_idx"])
    Fo = seg.compute_internal_force_2(fl, face_order=order)
    assert np.array_equal(Fo, ref)
    assert np.array_equal(Fo, flat.internal_forces(golden_io.snapshot_of(g), fl, g["vertex_bound"], order=order))
    bad = g["node_face_idx"].copy()
    if bad.size:
        bad[0] = g["face_nodes"].size // 3  # one past the face list
        with pytest.raises(Exception):
            seg.set_node_face_order(g["node_face_off"], bad)
    seg.set_node_face_order(None, None)
    assert np.array_equal(seg.compute_internal_force_2(fl), F)  # dropped: ascending order again
    seg.set_node_face_order(*order)
    # get_node_displacement (DEMO/segment_function.cpp:581-584)
    Fe = seg.compute_external_force(mean=g["mean"])
    d = seg.get_node_displacements(0.5)
    assert np.array_equal(d, (g["forces"] + g["meta"]["alpha"] * F) * 0.5)
    # flags decide the projection: with every node marked crossing nothing is projected
    F3 = seg.compute_internal_force_2(fl | 2)
    assert np.array_equal(F3[unprojected], F[unprojected]) and not np.array_equal(F3, F)


def test_points_in_tets_bit_exact(case):
    g, img, seg = case
    assert np.array_equal(seg.points_in_tets(g["pit_tet"], g["pit_pts"]), g["pit_inside"])


def test_full_evaluation_through_host_buffers(case, pkg):
    g, img, seg = case
    s = golden_io.snapshot_of(g)
    mean, F = seg.image_force_eval(s, mean_mode=pkg.MEAN_ZRAY, force_mode=pkg.FORCE_GATHER)
    np.testing.assert_allclose(mean, g["mean"], rtol=RTOL)
    assert np.array_equal(seg._phase_volume, g["phase_volume"])
    scale = np.abs(g["forces"]).max()
    np.testing.assert_allclose(F, g["forces"], rtol=1e-5, atol=1e-6 * scale)  # means differ in the last bits -> forces too
    _, F_pin = seg.image_force_eval(s, mean_mode=pkg.MEAN_ZRAY, force_mode=pkg.FORCE_GATHER, want_forces="pinned")
    assert np.array_equal(np.array(F_pin), F)  # zero-copy view of the pinned result buffer
    mean_t, F_t = seg.image_force_eval(s, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_ATOMIC)
    P = g["meta"]["nb_phase"]
    tot = np.array([g["tet_total"][g["tet_label"] == k + 1].sum() for k in range(P)])
    vol = np.array([g["tet_vol"][g["tet_label"] == k + 1].sum() for k in range(P)])
    np.testing.assert_allclose(mean_t, tot / vol, rtol=RTOL)


def test_partition_sums_compose(case, pkg):
    """Multi-GPU contract: partial results of disjoint partitions add up to the whole (emulated on one GPU)."""
    g, img, seg = case
    whole = seg.tet_integrate(per_tet=False)
    F_whole = seg.compute_external_force(mean=g["mean"])
    nt, nf, nn = seg.n_tets, seg.n_faces, seg.n_nodes_buffer
    tot = np.zeros_like(whole["phase_total"])
    vol = np.zeros_like(tot)
    F = np.zeros_like(F_whole)
    Fa = np.zeros_like(F_whole)
    R = 3
    for r in range(R):
        rng = lambda n: (n * r // R, n * (r + 1) // R)
        seg.set_partition(rng(nt), rng(nf), rng(nn))
        o = seg.tet_integrate(per_tet=False)
        tot += o["phase_total"]
        vol += o["phase_volume"]
        F += seg.compute_external_force(mode=pkg.FORCE_GATHER, mean=g["mean"])
        Fa += seg.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=g["mean"])
    seg.set_partition((0, nt), (0, nf), (0, nn))
    np.testing.assert_allclose(tot, whole["phase_total"], rtol=1e-12)
    np.testing.assert_allclose(vol, whole["phase_volume"], rtol=1e-12)
    # gather with a face partition: every rank adds its faces in order; the cross-rank sum re-associates (1e-12), it is
    # bit-exact only when each context holds all faces (node-only partition, checked below)
    np.testing.assert_allclose(F, F_whole, rtol=1e-12, atol=1e-12 * np.abs(F_whole).max())
    Fn = np.zeros_like(F_whole)
    for r in range(R):
        seg.set_partition((0, nt), (0, nf), (nn * r // R, nn * (r + 1) // R))
        Fn += seg.compute_external_force(mode=pkg.FORCE_GATHER, mean=g["mean"])
    seg.set_partition((0, nt), (0, nf), (0, nn))
    assert np.array_equal(Fn, F_whole)
    np.testing.assert_allclose(Fa, F_whole, rtol=RTOL, atol=RTOL * np.abs(F_whole).max())


# ---- seeded synthetic inputs against the flat oracle ------------------------------------------------

@pytest.mark.parametrize("dtype,n,edge,jitter", [(np.float32, 40, 3.1, 0.3), (np.uint16, 36, 5.0, 0.25), (np.uint8, 33, 7.7, 0.3),
                                                  (np.float64, 24, 12.0, 0.2), (np.float32, 30, 40.0, 0.1)])
def test_synthetic_vs_oracle(pkg, flat, dtype, n, edge, jitter):
    syn = pkg.synthetic
    vol = syn.shells_phantom(n, sigma=0.04, seed=7, dtype=np.float32)
    if dtype == np.float64:
        vol = vol.astype(np.float64) * 1.000001
    elif dtype != np.float32:
        vol = syn.shells_phantom(n, sigma=0.04, seed=7, dtype=dtype)
    P = 3
    snap = syn.build_case(vol, edge, P, (0.3, 0.7),
                          lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"], jitter=jitter, seed=11)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, P, 0.01)
    seg.set_snapshot(snap)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], c=[0.2, 0.5, 0.8])
    h, ns = seg.tet_sample_hash()
    assert np.array_equal(h, ref["hash"]) and np.array_equal(ns, ref["nsamp"])
    o = seg.tet_integrate(c=[0.2, 0.5, 0.8], include_bound=True)
    assert np.array_equal(o["vol"], ref["vol"])
    np.testing.assert_allclose(o["total"], ref["total"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o["variation"], ref["ad"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o["energy"], ref["sq"], rtol=RTOL, atol=1e-12)
    if len(snap["face_nodes"]):
        z = flat.zray(vol, snap, P, mode=0)
        mean = seg.update_average_intensity()
        assert np.array_equal(seg._phase_volume, z["count"])
        ok = z["count"] > 0
        np.testing.assert_allclose(seg._total_intensities[ok], z["total"][ok], rtol=RTOL)
        m = np.where(ok, z["mean"], 0.5)
        F_ref = flat.face_forces(vol, snap, P, m)
        assert np.array_equal(seg.compute_external_force(mean=m), F_ref)
        e = flat.zray(vol, snap, P, mode=1, mean=m, alpha=0.01)
        d, a = seg.compute_energy(mean=m)
        np.testing.assert_allclose([d, a], [e["energy_data"], e["energy_area"]], rtol=RTOL)
    img.close()


def test_edge_cases(pkg):
    vol = np.zeros((8, 8, 8), np.float32)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 2)
    with pytest.raises(pkg.DscGpuError):  # no snapshot yet
        seg.tet_integrate()
    empty = {"pos": np.zeros((0, 3)), "tet_nodes": np.zeros((0, 4), np.uint32), "tet_label": np.zeros(0, np.int32),
             "face_nodes": np.zeros((0, 3), np.uint32), "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)}
    seg.set_snapshot(empty)  # empty mesh is legal
    o = seg.tet_integrate()
    assert o["total"].size == 0 and np.all(o["phase_total"] == 0)
    assert seg.compute_external_force(mean=[0.1, 0.9]).shape == (0, 3)
    bad = dict(empty)
    bad["pos"] = np.zeros((4, 3))
    bad["tet_nodes"] = np.array([[0, 1, 2, 9]], np.uint32)  # node key out of range
    bad["tet_label"] = np.array([1], np.int32)
    with pytest.raises(pkg.DscGpuError):
        seg.set_snapshot(bad)
    # a degenerate (zero volume) tet and a tet fully outside the image must not crash and read zeros
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [-50, -50, -50], [-40, -50, -50], [-50, -40, -50], [-50, -50, -40]], float)
    snap = dict(empty)
    snap["pos"] = pos
    snap["tet_nodes"] = np.array([[0, 1, 2, 3], [4, 5, 6, 7]], np.uint32)
    snap["tet_label"] = np.array([1, 2], np.int32)
    seg.set_snapshot(snap)
    o = seg.tet_integrate()
    assert o["vol"][0] == 0.0 and o["total"][1] == 0.0 and abs(o["vol"][1] - 1000.0 / 6.0) < 1e-9
    img.close()


def test_rays_with_many_hits_take_the_slow_path(pkg, flat):
    """More stacked interfaces over one column than DSCGPU_MAX_RAY_HITS (the reference has no limit): such rays are sorted in place
    in the hit buffer; counts bit for bit as for every other ray."""
    n_layers = 121  # 120 interfaces per column: an even number of hits (rays with an odd number are dropped by the reference)
    rng = np.random.default_rng(5)
    vol = rng.uniform(0, 1, (n_layers + 4, 4, 4)).astype(np.float32)
    pos = []
    for k in range(n_layers + 1):
        z = k * 1.0
        pos += [[-4.0, -4.0, z], [12.0, -4.0, z], [-4.0, 12.0, z]]
    pos = np.array(pos, float)
    tets, labels = [], []
    for k in range(n_layers):
        a = 3 * k
        # a prism split in 3 tets between layer k and k+1
        for t in ([a, a + 1, a + 2, a + 3], [a + 1, a + 2, a + 3, a + 4], [a + 2, a + 3, a + 4, a + 5]):
            tets.append(t)
            labels.append(1 + (k % 2))
    tets = np.array(tets, np.uint32)
    labels = np.array(labels, np.int32)
    fn, ft, fb = pkg.synthetic.interface_faces(tets, labels, len(pos))
    snap = {"pos": pos, "tet_nodes": tets, "tet_label": labels, "face_nodes": fn, "face_tets": ft, "face_is_boundary": fb}
    z = flat.zray(vol, snap, 2, mode=0)
    assert z["hits_hist"][63] > 0 and z["count"].sum() > 0  # every column inside the prism crosses 120 stacked interface triangles
    img = pkg.Image3D(vol)
    img.build_sum_table()
    seg = pkg.SegmentFunction(img, 2)
    seg.set_snapshot(snap)
    for _ in range(2):  # the second pass runs with the hit buffer sized by the first (no host round trip inside)
        m = seg.update_average_intensity()
        assert np.array_equal(seg._phase_volume, z["count"])
        ok = z["count"] > 0
        np.testing.assert_allclose(seg._total_intensities[ok], z["total"][ok], rtol=RTOL)
    img.close()


@pytest.mark.parametrize("variant", [0, 128, 4224, 16512, 131072 | 16512])
def test_every_kernel_variant_classifies_identically(pkg, variant):
    """All implementations of the tet pass (incl. the FP32-filtered ones) must put every sample in the same voxel:
    per-tet sums of noisy data agree to rounding, and the volumes bit for bit."""
    for name in ("c1_e8_it2", "sph48_f32", "sh40_u16", "c1n_e9_j"):
        g = golden_io.load(name)
        img, seg = _seg(pkg, g)
        lab = g["tet_label"] != 0
        seg.set_tet_variant(variant)
        for lanes in ((8, 32) if variant == 0 else (0,)):
            seg.set_lanes_per_tet(lanes)
            o = seg.tet_integrate()
            assert np.array_equal(o["vol"][lab], g["tet_vol"][lab])
            np.testing.assert_allclose(o["total"][lab], g["tet_total"][lab], rtol=1e-12, atol=1e-13)
            P = g["meta"]["nb_phase"]
            tot = np.array([g["tet_total"][g["tet_label"] == k + 1].sum() for k in range(P)])
            np.testing.assert_allclose(o["phase_total"], tot, rtol=1e-9)
            o2 = seg.tet_integrate(per_tet=False)
            np.testing.assert_allclose(o2["phase_total"], tot, rtol=1e-9)
            np.testing.assert_allclose(o2["phase_volume"], o["phase_volume"], rtol=1e-12)
        img.close()


def test_filter_variants_on_large_interior_mesh(pkg, flat):
    """The FP32 filter only engages for tets strictly inside the volume; use a case where most tets are."""
    syn = pkg.synthetic
    vol = syn.shells_phantom(96, sigma=0.05, seed=3, dtype=np.float32)
    snap = syn.build_case(vol, 6.3, 3, (0.3, 0.7), lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"],
                          jitter=0.3, seed=17)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], want_hash=False)
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, 3)
    seg.set_snapshot(snap)
    for variant in (128, 4224, 16512, 131072 | 16512):
        seg.set_tet_variant(variant)
        o = seg.tet_integrate(include_bound=True)
        assert np.array_equal(o["vol"], ref["vol"])
        # float voxel sums of <= 729 terms are exact in double whatever the order: bit-equal totals <=> same voxels
        assert np.array_equal(o["total"], ref["total"]), variant
    img.close()


def test_relabel_energies(case):
    """get_energy_tetrahedron for every tet and candidate label, against the reference's own values."""
    g, img, seg = case
    e = seg.relabel_energies(g["tet_face_nodes"], g["tet_face_tets"], mean=g["mean"])
    lab = g["tet_label"] != 0
    np.testing.assert_allclose(e[lab], g["tet_relabel_energy"][lab], rtol=RTOL, atol=1e-12)
    assert np.all(e[~lab] == -1.0)
    # the decision relabel_tetrahedra takes from these numbers (Jacobi form: labels untouched) must agree too
    assert np.array_equal(np.argmin(e[lab], axis=1), np.argmin(g["tet_relabel_energy"][lab], axis=1)) or \
        np.allclose(np.sort(e[lab], axis=1)[:, 0], np.sort(g["tet_relabel_energy"][lab], axis=1)[:, 0], rtol=RTOL)


@pytest.mark.parametrize("dtype,P", [(np.float32, 6), (np.uint16, 8), (np.uint8, 5)])
def test_non_cubic_volume_and_many_phases(pkg, flat, dtype, P):
    """X != Y != Z and more than four phases (label-indexed lanes / accumulators up to DSCGPU_MAX_PHASE)."""
    syn = pkg.synthetic
    vol = np.ascontiguousarray(syn.shells_phantom(48, levels=(0.05, 0.35, 0.65, 0.95), radii=(0.45, 0.3, 0.15), sigma=0.08, seed=21, dtype=dtype)[5:41, 0:30, 3:48])
    assert vol.shape == (36, 30, 45)
    thr = tuple(np.linspace(0.1, 0.9, P - 1))
    snap = syn.build_case(vol, 5.5, P, thr, lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"], jitter=0.25, seed=3)
    assert len(np.unique(snap["tet_label"])) >= P  # every phase (and the boundary layer) is populated
    img = pkg.Image3D(vol)
    seg = pkg.SegmentFunction(img, P, 0.01)
    seg.set_snapshot(snap)
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], c=[0.0])
    h, ns = seg.tet_sample_hash()
    assert np.array_equal(h, ref["hash"]) and np.array_equal(ns, ref["nsamp"])
    lab = snap["tet_label"]
    for variant in (0, 128, 16512, 131072 | 16512):
        seg.set_tet_variant(variant)
        o = seg.tet_integrate()
        m = lab != 0
        assert np.array_equal(o["vol"][m], ref["vol"][m])
        np.testing.assert_allclose(o["total"][m], ref["total"][m], rtol=RTOL, atol=1e-12)
        tot, volu, mean = flat.phase_reduce(lab, ref["total"], ref["vol"], P)
        np.testing.assert_allclose(o["phase_total"], tot, rtol=RTOL)
        np.testing.assert_allclose(o["phase_volume"], volu, rtol=RTOL)
        sq, _, _ = flat.phase_reduce(lab, ref["sq"][:, 0], ref["vol"], P)
        np.testing.assert_allclose(o["phase_sumsq"], sq, rtol=RTOL)
    z = flat.zray(vol, snap, P, mode=0)
    seg.update_average_intensity()
    assert np.array_equal(seg._phase_volume, z["count"])
    ok = z["count"] > 0
    np.testing.assert_allclose(seg._total_intensities[ok], z["total"][ok], rtol=RTOL)
    mvals = np.where(ok, z["mean"], 0.5)
    assert np.array_equal(seg.compute_external_force(mean=mvals), flat.face_forces(vol, snap, P, mvals))
    np.testing.assert_allclose(seg.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=mvals), flat.face_forces(vol, snap, P, mvals), rtol=RTOL,
                               atol=RTOL * 10)
    e = flat.zray(vol, snap, P, mode=1, mean=mvals, alpha=0.01)
    d, a = seg.compute_energy(mean=mvals)
    np.testing.assert_allclose([d, a], [e["energy_data"], e["energy_area"]], rtol=RTOL)
    assert np.array_equal(img.sum_table(), flat.build_sum_table(vol))
    img.close()


def test_node_destinations_equal_the_restated_snapp_boundary(case):
    """SURVEY 8(f) row 3, second half: get_node_displacement (DEMO/segment_function.cpp:581-584) + the non-topological part of
    snapp_boundary (:466-541) against a numpy restatement of the same element-wise FP64 steps: bit for bit.  (The live reference is
    compared in tests/test_dropin_gpu.py.)"""
    g, img, seg = case
    rng = np.random.default_rng(11)
    n = len(g["pos"])
    F = rng.normal(0, 3.0, (n, 3))
    Fi = rng.normal(0, 1.0, (n, 3))
    flags = np.zeros(n, np.uint8)
    act = np.unique(g["face_nodes"])
    flags[act] = 1
    flags[act[::7]] |= 2
    vb = np.zeros(n, np.uint8)
    vb[act[::3]] = 1
    adapt = rng.uniform(0.5, 1.0, n)
    dims = np.array(g["volume"].shape[::-1], float)
    dt, alpha, thr, maxd = 0.7, seg.m_alpha, 2.4, 0.9
    dest, cur, st = seg.node_destinations(dt, thr, maxd, forces=F, internal_forces=Fi, node_flags=flags, vertex_bound=vb, dt_adapt=adapt)
    pos = g["pos"]
    length = lambda v: np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    dis = (F + alpha * Fi) * dt
    d = pos + dis
    b = vb.astype(bool)
    for k in range(3):
        dk = d[:, k].copy()
        dk[b & (pos[:, k] < thr)] = 0
        dk[b & (pos[:, k] > dims[k] - 1 - thr)] = dims[k] - 1
        dk[b & (dk < thr)] = 0
        dk[b & (dk > dims[k] - 1 - thr)] = dims[k] - 1
        d[:, k] = dk
    nd = d - pos
    stage1 = np.zeros(n, bool); stage1[act] = True
    ln = length(nd)
    max_dis = ln[stage1 & ~b].max()
    max_b = ln[stage1 & b].max()
    scale = 1.0
    if max_dis > maxd:
        scale = maxd / max_dis
        max_dis = maxd
    dis2 = nd * scale
    l2 = length(dis2)
    clip = b & (l2 > max_dis)
    dis2[clip] = dis2[clip] * (max_dis / l2[clip])[:, None]
    exp_dest = pos.copy()
    exp_dest[act] = pos[act] + dis2[act] * adapt[act][:, None]
    exp_cur = np.zeros((n, 3)); exp_cur[act] = dis2[act]
    assert np.array_equal(dest, exp_dest)
    assert np.array_equal(cur, exp_cur)
    assert st["max_boundary_dis"] == max_b and st["scale"] == scale
    assert st["real_max"] == (length(dis2) * adapt)[act].max()


def test_thickening_forces_and_face_statistics(case, flat):
    """A9b: the force loop of thickenning_surface (DEMO/segment_function.cpp:1990-2051) -- forces without the faces whose nodes are all
    vertex-bound, per-face mean and mean absolute deviation of the Gauss-point magnitudes -- against the oracle's restatement: bit for bit."""
    g, img, seg = case
    snap = golden_io.snapshot_of(g)
    vb = flat.vertex_bound(snap)
    F_ref, mu_ref, dev_ref, ev_ref = flat.thicken_forces(g["volume"], snap, g["meta"]["nb_phase"], g["mean"], vb)
    F, mu, dev, ev = seg.thicken_forces(mean=g["mean"], vertex_bound=vb)
    assert np.array_equal(ev, ev_ref) and 0 < ev.sum() <= len(ev)
    assert np.array_equal(F, F_ref)
    assert np.array_equal(mu, mu_ref) and np.array_equal(dev, dev_ref)
    F2, _, _, ev2 = seg.thicken_forces(mean=g["mean"])  # vertex-bound flags derived on the device
    assert np.array_equal(ev2, ev_ref) and np.array_equal(F2, F_ref)


def test_compact_force_read_back(case, pkg):
    """dscgpu_image_force_eval with the compact read-back: the rows of the nodes with interface faces, every other row of the full
    result is zero and the rows equal the full result's."""
    g, img, seg = case
    snap = golden_io.snapshot_of(g)
    mean, F = seg.image_force_eval(snap, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER)
    mean2, (keys, rows) = seg.image_force_eval(None, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER, want_forces="compact")
    # (the first call sums the tets in the chunks of the pipelined upload, the second in one pass: identical integers with the staged
    # kernel; volumes it cannot take -- a row pitch that is not a multiple of 16 bytes -- use the w2 kernel, whose double sums re-associate)
    np.testing.assert_allclose(mean, mean2, rtol=1e-13)
    assert np.array_equal(keys, np.unique(snap["face_nodes"]))
    if np.array_equal(mean, mean2):
        assert np.array_equal(rows, F[keys])
    else:
        np.testing.assert_allclose(rows, F[keys], rtol=1e-10, atol=1e-12)
    rest = np.ones(len(F), bool); rest[keys] = False
    assert not F[rest].any()


@pytest.mark.parametrize("name", ["sph48_f32", "sh40_u16", "c1n_e9_j"])
def test_pipelined_upload_orders_give_the_same_result(pkg, monkeypatch, name):
    """dscgpu_image_force_eval with a host snapshot overlaps the upload with the tet pass in chunks (face arrays behind chunk 0, node -> face
    table built meanwhile, exact list worked off per chunk, shrinking chunks): every order and chunk count gives the numbers of the plain
    commit-then-evaluate path -- bit for bit where the staged kernel runs (integer sums), 1e-13 where the w2 kernel's double sums re-associate."""
    g = golden_io.load(name)
    snap = golden_io.snapshot_of(g)
    P = g["meta"]["nb_phase"]
    results = []
    for flags, chunks, ratio in [("0", "0", "1"), ("0", "4", "1"), ("3", "3", "1"), ("7", "4", "0.85"), ("7", "7", "0.5"), ("2", "16", "1")]:
        monkeypatch.setenv("DSCGPU_PIPELINE_FLAGS", flags)
        monkeypatch.setenv("DSCGPU_PIPELINE_CHUNKS", chunks)
        monkeypatch.setenv("DSCGPU_PIPELINE_RATIO", ratio)
        img = pkg.Image3D(g["volume"])
        seg = pkg.SegmentFunction(img, P)
        mean, F = seg.image_force_eval(snap, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER)
        mean_c, (keys, rows) = seg.image_force_eval(snap, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER, want_forces="compact")
        staged = seg.get_tet_variant() & 131072
        assert np.array_equal(mean, mean_c) and np.array_equal(rows, F[keys]) and np.array_equal(keys, np.unique(snap["face_nodes"]))
        results.append((mean, F))
        img.close()
    m0, F0 = results[0]
    for m, F in results[1:]:
        if staged and name != "c1n_e9_j":  # (a 100-byte row pitch cannot be staged by TMA: w2 kernel)
            assert np.array_equal(m, m0) and np.array_equal(F, F0)
        else:
            np.testing.assert_allclose(m, m0, rtol=1e-13)
            np.testing.assert_allclose(F, F0, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("variant", [0, 16512, 131072 | 16512])
def test_free_slots_of_a_key_indexed_snapshot_are_skipped(pkg, flat, variant):
    """Slots that hold no tet (label DSCGPU_LABEL_FREE = -1, nodes 0 0 0 0) take part in nothing, also with include_bound: the sums are
    those of the mesh without them, their per-tet outputs are -1."""
    g = golden_io.load("sph48_f32")
    snap = golden_io.snapshot_of(g)
    nt = len(snap["tet_label"])
    free = np.arange(5, nt + 40, 97)
    tn = np.insert(snap["tet_nodes"], np.minimum(free, nt), 0, axis=0)
    tl = np.insert(snap["tet_label"], np.minimum(free, nt), -1)
    # face_tets index the tet array: shift them past the inserted slots
    is_free = tl == -1
    new_index = np.flatnonzero(~is_free)
    ft = snap["face_tets"].copy()
    ok = ft != 0xFFFFFFFF
    ft[ok] = new_index[ft[ok]].astype(np.uint32)
    snap2 = dict(snap, tet_nodes=tn.astype(np.uint32), tet_label=tl.astype(np.int32), face_tets=ft)
    img = pkg.Image3D(g["volume"])
    seg = pkg.SegmentFunction(img, g["meta"]["nb_phase"])
    seg.set_tet_variant(variant)
    seg.set_snapshot(snap)
    a = seg.tet_integrate(include_bound=True)
    Fa = seg.compute_external_force(mean=a["phase_mean"])
    seg.set_snapshot(snap2)
    b = seg.tet_integrate(include_bound=True)
    assert np.all(b["total"][is_free] == -1.0) and np.all(b["vol"][is_free] == -1.0)
    assert np.array_equal(b["total"][~is_free], a["total"]) and np.array_equal(b["vol"][~is_free], a["vol"])
    np.testing.assert_allclose(b["phase_total"], a["phase_total"], rtol=1e-12)
    assert np.array_equal(seg.compute_external_force(mean=a["phase_mean"]), Fa)
    img.close()


def test_read_forces_compact_from_a_device_array(case, pkg):
    """dscgpu_read_forces_compact: the rows of the nodes with interface faces out of a device-resident force array."""
    torch = pytest.importorskip("torch")
    g, img, seg = case
    P = g["meta"]["nb_phase"]
    n = len(g["pos"])
    d_mean = torch.tensor(np.asarray(g["mean"], np.float64), device="cuda")
    d_F = torch.zeros(3 * n, dtype=torch.float64, device="cuda")
    seg.face_forces_dev(d_mean.data_ptr(), d_F.data_ptr(), pkg.FORCE_GATHER)
    keys, rows = seg.read_forces_compact(d_F.data_ptr())
    F = d_F.cpu().numpy().reshape(-1, 3)
    assert np.array_equal(keys, np.unique(g["face_nodes"])) and np.array_equal(rows, F[keys])
    assert np.array_equal(F, g["forces"].reshape(-1, 3)) or np.allclose(F, g["forces"].reshape(-1, 3), rtol=1e-6, atol=1e-9)
