"""world_size-2 gloo run of the multi-GPU control flow (partition + two all-reduces) on CPU.

The compute stand-in is the flat oracle restricted to the rank's partition -- this tests the host-side sharding logic
(distributed.partition / evaluate), not the kernels; the GPU parity tests cover dscgpu_set_partition itself."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_package
    from oracle import flat
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = golden_io.load("sph48_f32")
    s = golden_io.snapshot_of(g)
    P = g["meta"]["nb_phase"]
    (tb, te), (fb, fe), _ = pkg.distributed.partitions(len(s["tet_nodes"]), len(s["face_nodes"]), len(s["pos"]), world, rank)

    def phase_sums():
        mask = np.zeros(len(s["tet_nodes"]), np.uint8)
        mask[tb:te] = (s["tet_label"][tb:te] != 0)
        o = flat.tet_integrate(g["volume"], s["tet_nodes"], s["pos"], c=[0.0], tet_mask=mask, want_hash=False)
        lab = np.where(mask.astype(bool), s["tet_label"], 0)
        tot, vol, _ = flat.phase_reduce(lab, np.where(mask, o["total"], 0.0), np.where(mask, o["vol"], 0.0), P)
        sq, _, _ = flat.phase_reduce(lab, np.where(mask, o["sq"][:, 0], 0.0), np.where(mask, o["vol"], 0.0), P)
        return np.concatenate([tot, sq, vol])

    def forces(mean):
        part = dict(s)
        part["face_nodes"] = s["face_nodes"][fb:fe]
        part["face_tets"] = s["face_tets"][fb:fe]
        part["face_is_boundary"] = s["face_is_boundary"][fb:fe]
        return flat.face_forces(g["volume"], part, P, mean)

    def all_reduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    mean, F = pkg.distributed.evaluate(phase_sums, forces, P, all_reduce)

    # handle exchange of the peer-memory path (connect_peers) with a stand-in for the CUDA side: every rank must end up with
    # all windows in rank order and map exactly the other ranks'
    class FakeSeg:
        def __init__(self):
            self.connected = None

        def comm_create(self, r, w, cap):
            assert (r, w, cap) == (rank, world, 123)
            return bytes([r]) * 64

        def comm_connect(self, handles):
            self.connected = handles

    def all_gather_object(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    fake = FakeSeg()
    handles = pkg.distributed.connect_peers(fake, rank, world, 123, all_gather_object)
    assert handles == [bytes([r]) * 64 for r in range(world)] and fake.connected == handles
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), mean=mean, F=F)
    dist.destroy_process_group()


def test_partition_covers_everything(pkg):
    for n in (0, 1, 7, 1000, 1971054):
        for world in (1, 2, 3, 8):
            parts = [pkg.distributed.partition(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for a, b in zip(parts, parts[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pkg.distributed.partition(10, 2, 2)


def test_balanced_cuts(pkg):
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 1000, 100000):
        cost = rng.integers(0, 730, n).astype(float)
        cost[rng.uniform(size=n) < 0.1] = 0.0
        for world in (1, 2, 3, 8):
            cuts = pkg.distributed.balanced_cuts(cost, world)
            assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == n
            assert all(a <= b for a, b in zip(cuts, cuts[1:]))
            if n >= 1000:
                loads = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
                assert max(loads) - min(loads) <= 2 * 729 + 1e-9
    assert pkg.distributed.balanced_cuts(np.zeros(10), 2) == [0, 5, 10]


def test_two_rank_evaluation_equals_single_rank(pkg, flat, tmp_path):
    world = 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["mean"], r1["mean"]) and np.array_equal(r0["F"], r1["F"])
    # single-rank answer
    g = golden_io.load("sph48_f32")
    s = golden_io.snapshot_of(g)
    P = g["meta"]["nb_phase"]
    mask = (s["tet_label"] != 0).astype(np.uint8)
    o = flat.tet_integrate(g["volume"], s["tet_nodes"], s["pos"], tet_mask=mask, want_hash=False)
    tot, vol, mean = flat.phase_reduce(np.where(mask.astype(bool), s["tet_label"], 0), np.where(mask, o["total"], 0.0), np.where(mask, o["vol"], 0.0), P)
    np.testing.assert_allclose(r0["mean"], mean, rtol=1e-12)
    F = flat.face_forces(g["volume"], s, P, r0["mean"])
    np.testing.assert_allclose(r0["F"], F, rtol=1e-9, atol=1e-12 * np.abs(F).max())
