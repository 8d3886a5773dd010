"""The fused exchange kernels (csrc/dsc_comm.cuh) on one GPU: world = 1, and world = 2 with two contexts in one process whose
kernels wait for each other's flags on two streams.  The result must equal the single-context evaluation (sums to rounding,
forces to the 1e-6 of the atomic scatter), be identical on both ranks, and stay right over many epochs and a snapshot change."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _contexts(pkg, g, world):
    segs = []
    for r in range(world):
        img = pkg.Image3D(g["volume"])
        seg = pkg.SegmentFunction(img, g["meta"]["nb_phase"], g["meta"]["alpha"])
        segs.append((img, seg))
    return segs


def _set(pkg, segs, snap, world):
    nt, nf, nn = len(snap["tet_nodes"]), len(snap["face_nodes"]), len(snap["pos"])
    for r, (img, seg) in enumerate(segs):
        seg.set_snapshot(snap)
        seg.set_partition(*pkg.distributed.partitions(nt, nf, nn, world, r))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_exchange_equals_single_context(pkg, world):
    torch = pytest.importorskip("torch")
    g = golden_io.load("c1_e6")
    snap = golden_io.snapshot_of(g)
    P = g["meta"]["nb_phase"]
    nn = len(snap["pos"])
    ref_img = pkg.Image3D(g["volume"])
    ref = pkg.SegmentFunction(ref_img, P, g["meta"]["alpha"])
    ref.set_snapshot(snap)
    o = ref.tet_integrate(per_tet=False)
    F_ref = ref.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=o["phase_mean"])

    segs = _contexts(pkg, g, world)
    _set(pkg, segs, snap, world)
    for r, (img, seg) in enumerate(segs):
        seg.comm_create(r, world, cap_rows=nn + 100)
    wins = [seg.comm_window() for _, seg in segs]
    for _, seg in segs:
        seg.comm_connect_local(wins)
        seg.compute_external_force(mean=seg.tet_integrate(per_tet=False)["phase_mean"])  # warm-up: one-time synchronising set-up (allocations) must not sit between the exchange launches
    bufs = [(torch.zeros(3 * P, dtype=torch.float64, device="cuda"), torch.zeros(P, dtype=torch.float64, device="cuda"),
             torch.full((3 * nn,), 7.0, dtype=torch.float64, device="cuda")) for _ in range(world)]
    for epoch in range(5):  # both halves of the windows, several times
        for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
            seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
        for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
            seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_F.data_ptr())
        for img, seg in segs:
            seg.comm_error()
        outs = [(s_.cpu().numpy().copy(), m_.cpu().numpy().copy(), f_.cpu().numpy().reshape(-1, 3).copy()) for s_, m_, f_ in bufs]
        for s_, m_, f_ in outs[1:]:
            assert np.array_equal(s_, outs[0][0]) and np.array_equal(m_, outs[0][1])  # peers are added in rank order everywhere
            np.testing.assert_allclose(f_, outs[0][2], rtol=1e-12, atol=1e-12 * np.abs(F_ref).max())
        np.testing.assert_allclose(outs[0][0][:P], o["phase_total"], rtol=1e-12)
        np.testing.assert_allclose(outs[0][0][2 * P:], o["phase_volume"], rtol=1e-12)
        np.testing.assert_allclose(outs[0][1], o["phase_mean"], rtol=1e-12)
        np.testing.assert_allclose(outs[0][2], F_ref, rtol=RTOL, atol=RTOL * np.abs(F_ref).max())
    # a different snapshot with fewer interface faces: rows left over from the larger one must not leak into the result
    g2 = golden_io.load("c1_e8_it2")
    if g2["volume"].shape == g["volume"].shape and g2["meta"]["nb_phase"] == P and len(golden_io.snapshot_of(g2)["pos"]) <= nn + 100:
        snap2 = golden_io.snapshot_of(g2)
        ref.set_snapshot(snap2)
        o2 = ref.tet_integrate(per_tet=False)
        F2 = ref.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=o2["phase_mean"])
        _set(pkg, segs, snap2, world)
        nn2 = len(snap2["pos"])
        bufs2 = [(torch.zeros(3 * P, dtype=torch.float64, device="cuda"), torch.zeros(P, dtype=torch.float64, device="cuda"),
                  torch.zeros(3 * nn2, dtype=torch.float64, device="cuda")) for _ in range(world)]
        for epoch in range(3):
            for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs2):
                seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
            for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs2):
                seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_F.data_ptr())
            for img, seg in segs:
                seg.comm_error()
            f0 = bufs2[0][2].cpu().numpy().reshape(-1, 3)
            np.testing.assert_allclose(bufs2[0][1].cpu().numpy(), o2["phase_mean"], rtol=1e-12)
            np.testing.assert_allclose(f0, F2, rtol=RTOL, atol=RTOL * np.abs(F2).max())
    for img, seg in segs:
        img.close()
    ref_img.close()


def test_exchange_without_window_fails_loudly(pkg):
    torch = pytest.importorskip("torch")
    g = golden_io.load("sph48_f32")
    img = pkg.Image3D(g["volume"])
    seg = pkg.SegmentFunction(img, g["meta"]["nb_phase"])
    seg.set_snapshot(golden_io.snapshot_of(g))
    d = torch.zeros(64, dtype=torch.float64, device="cuda")
    with pytest.raises(pkg.DscGpuError):
        seg.tet_phase_sums_allreduce_dev(d.data_ptr(), d.data_ptr())
    img.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_commit_and_incremental_update_under_a_partition(pkg, world):
    """dscgpu_commit_snapshot_sharded: every rank uploads 1/world of the staged bytes and receives the rest through the windows;
    the assembled snapshot must be the staged one on every rank (same evaluation as a plain commit).  Then an incremental
    update that replaces the face list under the partition: the ranges must be set again (a stale partition is an error, not
    forces added world times), after which the exchange equals one context evaluating the updated snapshot."""
    torch = pytest.importorskip("torch")
    g = golden_io.load("c1_e6")
    snap = golden_io.snapshot_of(g)
    P = g["meta"]["nb_phase"]
    nn, nt, nf = len(snap["pos"]), len(snap["tet_nodes"]), len(snap["face_nodes"])
    ref_img = pkg.Image3D(g["volume"])
    ref = pkg.SegmentFunction(ref_img, P, g["meta"]["alpha"])
    ref.set_snapshot(snap)
    o = ref.tet_integrate(per_tet=False)
    F_ref = ref.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=o["phase_mean"])
    segs = _contexts(pkg, g, world)
    for r, (img, seg) in enumerate(segs):
        seg.comm_create(r, world, cap_rows=nn + 100, snapshot_bytes=32 * nn + 20 * nt + 21 * nf + 4096)
    wins = [seg.comm_window() for _, seg in segs]
    for _, seg in segs:
        seg.comm_connect_local(wins)
        # one process, one GPU: a cudaMalloc inside the collective would wait for the peers' spinning kernels -- size every buffer first
        seg.set_snapshot(snap)
        seg.compute_external_force(mean=seg.tet_integrate(per_tet=False)["phase_mean"])
    for rep in range(3):  # several epochs of the commit protocol (acknowledgements)
        for _, seg in segs:
            st = seg.staging(nn, nt, nf)
            st["pos4"][:, :3] = snap["pos"]
            st["pos4"][:, 3] = 0
            for k in ("tet_nodes", "tet_label", "face_nodes", "face_tets", "face_is_boundary"):
                st[k][...] = snap[k]
        for _, seg in segs:
            seg.commit_snapshot_sharded()
        for img, seg in segs:
            seg.comm_error()  # (synchronises) the commit protocol itself must not have timed out
        for r, (img, seg) in enumerate(segs):
            seg.set_partition(*pkg.distributed.partitions(nt, nf, nn, world, r))
        bufs = [(torch.zeros(3 * P, dtype=torch.float64, device="cuda"), torch.zeros(P, dtype=torch.float64, device="cuda"),
                 torch.zeros(3 * nn, dtype=torch.float64, device="cuda")) for _ in range(world)]
        for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
            seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
        for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
            seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_F.data_ptr())
        for img, seg in segs:
            seg.comm_error()
        for d_sums, d_mean, d_F in bufs:
            np.testing.assert_allclose(d_mean.cpu().numpy(), o["phase_mean"], rtol=1e-12)
            np.testing.assert_allclose(d_F.cpu().numpy().reshape(-1, 3), F_ref, rtol=RTOL, atol=RTOL * np.abs(F_ref).max())
    # incremental update: drop the last 10 interface faces, move a few nodes
    keep = nf - 10
    rng = np.random.default_rng(3)
    keys = np.sort(rng.choice(nn, 50, replace=False)).astype(np.uint32)
    pos2 = snap["pos"].copy()
    pos2[keys] += rng.uniform(-0.05, 0.05, (50, 3))
    snap2 = dict(snap, pos=pos2, face_nodes=snap["face_nodes"][:keep], face_tets=snap["face_tets"][:keep], face_is_boundary=snap["face_is_boundary"][:keep])
    ref.set_snapshot(snap2)
    o2 = ref.tet_integrate(per_tet=False)
    F2 = ref.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=o2["phase_mean"])
    for _, seg in segs:
        seg.update_snapshot(node_keys=keys, pos=pos2[keys], faces={k: snap2[k] for k in ("face_nodes", "face_tets", "face_is_boundary")})
    d = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
    with pytest.raises(pkg.DscGpuError, match="dscgpu_set_partition"):
        segs[0][1].face_forces_dev(d.data_ptr(), d.data_ptr(), pkg.FORCE_ATOMIC)
    for r, (img, seg) in enumerate(segs):
        seg.set_partition(*pkg.distributed.partitions(nt, keep, nn, world, r))
    for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
        seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
    for (img, seg), (d_sums, d_mean, d_F) in zip(segs, bufs):
        seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_F.data_ptr())
    for img, seg in segs:
        seg.comm_error()
    for d_sums, d_mean, d_F in bufs:
        np.testing.assert_allclose(d_mean.cpu().numpy(), o2["phase_mean"], rtol=1e-12)
        np.testing.assert_allclose(d_F.cpu().numpy().reshape(-1, 3), F2, rtol=RTOL, atol=RTOL * np.abs(F2).max())
    for img, seg in segs:
        img.close()
    ref_img.close()
