"""dscgpu_update_snapshot.  A context that received snapshot S0 in full and then the difference to S1
(moved nodes, relabelled / re-wired tets, the new interface-face list) must give exactly what a fresh context given S1 gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _outputs(pkg, seg, P):
    o = seg.tet_integrate()
    mean = seg.update_average_intensity().copy()
    vol = seg._phase_volume.copy()
    ok = vol > 0
    m = np.where(ok, mean, 0.5)
    F = seg.compute_external_force(mean=m)
    d, a = seg.compute_energy(mean=m)
    return o["total"], o["vol"], o["phase_total"], mean, vol, F, d, a


def test_update_equals_full_snapshot(pkg, flat):
    syn = pkg.synthetic
    vol = syn.shells_phantom(48, sigma=0.05, seed=9, dtype=np.float32)
    P = 3
    mean_fn = lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False)["mean"]
    s0 = syn.build_case(vol, 5.7, P, (0.3, 0.7), mean_fn, jitter=0.25, seed=5)
    rng = np.random.default_rng(1)
    n_nodes, n_tets = len(s0["pos"]), len(s0["tet_nodes"])
    # S1: 40 % of the nodes move a little, 4 % of the labelled tets change label, a few tets get their nodes permuted
    moved = np.flatnonzero(rng.uniform(size=n_nodes) < 0.4).astype(np.uint32)
    pos1 = s0["pos"].copy()
    pos1[moved] += rng.uniform(-0.3, 0.3, (len(moved), 3))
    lab1 = s0["tet_label"].copy()
    cand = np.flatnonzero(lab1 != 0)
    relab = rng.choice(cand, size=max(1, len(cand) // 25), replace=False)
    lab1[relab] = 1 + (lab1[relab] % P)
    tn1 = s0["tet_nodes"].copy()
    perm = rng.choice(n_tets, size=50, replace=False)
    tn1[perm] = tn1[perm][:, [1, 2, 0, 3]]  # an even permutation of the first three nodes: same tet, other get_nodes order
    changed = np.unique(np.concatenate([relab, perm])).astype(np.uint32)
    fn, ft, fb = syn.interface_faces(tn1, lab1, n_nodes)
    s1 = {"pos": pos1, "tet_nodes": tn1, "tet_label": lab1, "face_nodes": fn, "face_tets": ft, "face_is_boundary": fb}
    assert len(fn) != len(s0["face_nodes"])

    img_a = pkg.Image3D(vol)
    seg_a = pkg.SegmentFunction(img_a, P, 0.01)
    seg_a.set_snapshot(s0)
    _outputs(pkg, seg_a, P)  # state of a previous evaluation lying around
    seg_a.update_snapshot(node_keys=moved, pos=pos1[moved], tet_index=changed, tet_nodes=tn1[changed], tet_label=lab1[changed],
                          faces={"face_nodes": fn, "face_tets": ft, "face_is_boundary": fb})
    got = _outputs(pkg, seg_a, P)

    img_b = pkg.Image3D(vol)
    seg_b = pkg.SegmentFunction(img_b, P, 0.01)
    seg_b.set_snapshot(s1)
    want = _outputs(pkg, seg_b, P)
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # the fused evaluation must use the resident (updated) snapshot, not the stale pinned staging of S0
    mean_a, F_a = seg_a.image_force_eval(None, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER)
    mean_b, F_b = seg_b.image_force_eval(s1, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER)
    # (a full snapshot is uploaded in chunks with one tet launch per chunk: other partial-sum boundaries, last-bit differences)
    np.testing.assert_allclose(mean_a, mean_b, rtol=1e-12)
    np.testing.assert_allclose(F_a, F_b, rtol=1e-9, atol=1e-12 * np.abs(F_b).max())
    # all positions at once (keys omitted), faces kept
    seg_a.update_snapshot(pos=s0["pos"])
    seg_b.set_snapshot(dict(s1, pos=s0["pos"]))
    assert np.array_equal(seg_a.tet_integrate()["total"], seg_b.tet_integrate()["total"])
    assert np.array_equal(seg_a.compute_external_force(mean=[0.1, 0.5, 0.9]), seg_b.compute_external_force(mean=[0.1, 0.5, 0.9]))
    # zero-copy staging: fill the pinned arrays, commit (twice: the staged update may be committed again)
    st = seg_a.update_staging(len(moved), True, len(changed), len(fn))
    st["node_keys"][...] = moved
    st["pos"][...] = pos1[moved]
    st["tet_index"][...] = changed
    st["tet_nodes"][...] = tn1[changed]
    st["tet_label"][...] = lab1[changed]
    st["face_nodes"][...] = fn
    st["face_tets"][...] = ft
    st["face_is_boundary"][...] = fb
    seg_a.update_commit(True)
    seg_a.update_commit(True)
    seg_b.set_snapshot(s1)
    for g, w in zip(_outputs(pkg, seg_a, P), _outputs(pkg, seg_b, P)):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # bad keys are rejected loudly
    seg_a.update_snapshot(node_keys=np.array([n_nodes + 5], np.uint32), pos=np.zeros((1, 3)))
    with pytest.raises(pkg.DscGpuError):
        seg_a.tet_integrate()
    img_a.close()
    img_b.close()
