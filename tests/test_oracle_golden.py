"""Pins the flat CPU oracle (oracle/flat_oracle.cpp) against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py -> oracle/_ref/dsc_ref).  Everything is bit-exact: the
oracle restates the reference's operation order and is built without FP contraction."""
import numpy as np
import pytest

import golden_io


@pytest.fixture(scope="module", params=golden_io.CASES)
def case(request):
    g = golden_io.load(request.param)
    return g, golden_io.snapshot_of(g)


def test_known_answers_c1():
    # the survey's independent probe values for BASELINE config #1
    g = golden_io.load("c1_e6")
    assert g["meta"]["n_tets"] == 41154 and g["meta"]["n_if"] == 7281 and g["meta"]["n_nodes_buffer"] == 8000
    np.testing.assert_allclose(g["mean"], [0.00707454434289, 0.210471379735, 0.67494909049], rtol=1e-11)
    assert g["phase_volume"].tolist() == [505936.0, 344444.0, 145173.0]
    np.testing.assert_allclose(g["total"], [3579.26666667, 72495.6039216, 97984.3843137], rtol=1e-11)
    np.testing.assert_allclose(np.linalg.norm(g["forces"]), 235.430971862, rtol=1e-11)
    np.testing.assert_allclose(np.abs(g["forces"]).sum() * 0 + np.linalg.norm(g["forces"], axis=1).sum(), 9811.82756599, rtol=1e-11)
    np.testing.assert_allclose(g["forces"].sum(axis=0), [-2161.08049645, -2193.72204026, -1632.64253595], rtol=1e-11)
    np.testing.assert_allclose(g["forces"][421], [-2.6454491454, -2.6454491454, -1.3227245727], rtol=1e-10)
    np.testing.assert_allclose([g["meta"]["energy_data"], g["meta"]["energy_area"]], [4661.39, 83.1473], rtol=1e-5)


def test_tet_sampling_bit_exact(case, flat):
    g, s = case
    lab = g["tet_label"] != 0
    o = flat.tet_integrate(g["volume"], s["tet_nodes"], s["pos"], c=g["mean"])
    assert np.array_equal(o["hash"], g["tet_hash"])       # every sample lands in the reference's voxel
    assert np.array_equal(o["nsamp"], g["tet_nsamp"])     # sample level
    for mine, ref in (("total", "tet_total"), ("vol", "tet_vol"), ("mean", "tet_mean"), ("ad", "tet_variation"), ("sq", "tet_energy")):
        assert np.array_equal(o[mine][lab], g[ref][lab]), mine


def test_zray_means_bit_exact(case, flat):
    g, s = case
    P = g["meta"]["nb_phase"]
    z = flat.zray(g["volume"], s, P, mode=0)
    assert np.array_equal(z["count"], g["phase_volume"])
    assert np.array_equal(z["total"], g["total"])
    assert np.array_equal(z["mean"], g["mean"])


def test_forces_bit_exact(case, flat):
    g, s = case
    F = flat.face_forces(g["volume"], s, g["meta"]["nb_phase"], g["mean"])
    assert np.array_equal(F, g["forces"])


def test_energy_and_vertex_bound(case, flat):
    g, s = case
    vb = flat.vertex_bound(s)
    assert np.array_equal(vb, g["vertex_bound"])
    e = flat.zray(g["volume"], s, g["meta"]["nb_phase"], mode=1, mean=g["mean"], alpha=g["meta"]["alpha"])
    assert e["energy_data"] == g["meta"]["energy_data"]
    assert e["energy_area"] == g["meta"]["energy_area"]


def test_points_in_tets(case, flat):
    g, s = case
    assert np.array_equal(flat.points_in_tets(g["pit_tet"], g["pit_pts"], s["tet_nodes"], s["pos"]), g["pit_inside"])


def test_sum_table_matches_cumsum_order(flat):
    g = golden_io.load("sh40_u16")
    t = flat.build_sum_table(g["volume"])
    v = flat.decode(g["volume"])
    ref = np.empty_like(v)
    ref[0] = v[0]
    for z in range(1, v.shape[0]):
        ref[z] = ref[z - 1] + v[z]
    assert np.array_equal(t, ref)


def test_multithreaded_oracle_matches_single_thread(flat):
    g = golden_io.load("c1n_e9_j")
    s = golden_io.snapshot_of(g)
    a = flat.tet_integrate(g["volume"], s["tet_nodes"], s["pos"], threads=1)
    b = flat.tet_integrate(g["volume"], s["tet_nodes"], s["pos"], threads=4)
    assert np.array_equal(a["total"], b["total"]) and np.array_equal(a["hash"], b["hash"])
    F1 = flat.face_forces(g["volume"], s, 3, g["mean"], threads=1)
    F4 = flat.face_forces(g["volume"], s, 3, g["mean"], threads=4)
    np.testing.assert_allclose(F4, F1, rtol=1e-9, atol=1e-12)


def test_relabel_energies_bit_exact(case, flat):
    g, s = case
    e = flat.relabel_energies(g["volume"], s, g["tet_face_nodes"], g["tet_face_tets"], g["meta"]["nb_phase"], g["mean"], g["meta"]["alpha"])
    lab = g["tet_label"] != 0
    assert np.array_equal(e[lab], g["tet_relabel_energy"][lab])


def test_internal_forces_match_the_reference(flat):
    """compute_internal_force_2.  With the get_faces(node) order of the reference's complex (node_face_off /
    node_face_idx of the fixture: what DSC::get_normal(node) iterates) every node is bit-identical.  Without it the nodes that are
    not projected on their normal (crossing or not interface) are bit-identical and the projected ones differ by rounding."""
    for name in golden_io.CASES:
        g = golden_io.load(name)
        s = golden_io.snapshot_of(g)
        F = flat.internal_forces(s, g["node_flags"], g["vertex_bound"])
        ref = g["internal_forces"]
        fl = g["node_flags"]
        unprojected = ((fl & 1) == 0) | ((fl & 2) != 0)
        assert np.array_equal(F[unprojected], ref[unprojected])
        np.testing.assert_allclose(F, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
        assert np.abs(ref).max() > 1.0  # the fixture exercises the path
        assert np.array_equal(flat.internal_forces(s, fl), F)  # vertex_bound derived from the snapshot
        off, idx = g["node_face_off"], g["node_face_idx"]
        assert off.size == fl.size + 1 and off[-1] == idx.size and idx.size > 0
        assert np.array_equal(flat.internal_forces(s, fl, g["vertex_bound"], order=(off, idx)), ref)
        # the order lists exactly the interface faces around each projected node
        cnt = np.bincount(s["face_nodes"].reshape(-1), minlength=fl.size)
        projected = ~unprojected
        assert np.array_equal(np.diff(off)[projected], cnt[projected])


def test_initialisation_pass_matches_the_reference(flat):
    """histogram of per-tet means, label rule.  c1_e6 carries the labels the reference's own
    initialization_discrete_opt assigned (labels=otsu, no iteration): they pin histogram, thresholds and rule together."""
    for name in golden_io.CASES:
        g = golden_io.load(name)
        lab = g["tet_label"]
        assert np.array_equal(flat.init_histogram(lab, g["tet_mean"]), g["init_hist"])
        assert np.array_equal(flat.init_labels(lab, g["tet_mean"], g["init_thres"]), g["init_labels"])
    g = golden_io.load("c1_e6")
    assert np.array_equal(g["init_labels"], g["tet_label"])
