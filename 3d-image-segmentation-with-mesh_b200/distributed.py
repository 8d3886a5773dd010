"""Multi-GPU composition of one image-force evaluation.

One process per GPU.  The volume and the mesh snapshot are replicated; the work is partitioned into contiguous ranges of
tets (per-tet integration), interface faces (scatter force kernel) and nodes (gather force kernel).  The only exchanges
are two all-reduces: the 3*nb_phase per-phase sums (the means are needed by the force pass) and the 3*n_nodes_buffer
vertex-force array.  `evaluate` is written against two callables so that the same control flow runs with the CUDA
library over NCCL on GPUs and with any stand-in compute over gloo in the CPU tests.
"""
import numpy as np


def partition(n, world, rank):
    """Contiguous range [begin, end) of rank `rank` when n items are split over `world` ranks (sizes differ by at most 1)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return n * rank // world, n * (rank + 1) // world


def balanced_cuts(cost, world):
    """Cut points [0, c1, ..., n] of `world` contiguous ranges with (nearly) equal total cost: range r is [cuts[r], cuts[r+1]).
    `cost` is a non-negative array with one entry per item (e.g. the sample count of a tet, ~0 for label-0 tets)."""
    cost = np.asarray(cost, np.float64)
    n = len(cost)
    if world < 1:
        raise ValueError("bad world")
    total = cost.sum()
    if n == 0 or total <= 0:
        return [n * r // world for r in range(world + 1)]
    cum = np.cumsum(cost)
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(cum, total * r / world, side="left")) + 1
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return cuts


def partitions(n_tets, n_faces, n_nodes, world, rank):
    return partition(n_tets, world, rank), partition(n_faces, world, rank), partition(n_nodes, world, rank)


def evaluate(phase_sums_fn, forces_fn, nb_phase, all_reduce):
    """One evaluation on this rank's partition.

    phase_sums_fn() -> array [3*nb_phase] (total, sumsq, volume of this rank's tets)
    forces_fn(mean) -> array [n_nodes_buffer, 3] holding this rank's contribution (zeros elsewhere)
    all_reduce(x)   -> sums x over ranks in place (NCCL / gloo) and returns it
    Returns (mean[nb_phase], forces) identical on every rank.
    """
    sums = all_reduce(phase_sums_fn())
    total, volume = sums[:nb_phase], sums[2 * nb_phase:3 * nb_phase]
    mean = total / volume
    forces = all_reduce(forces_fn(mean))
    return mean, forces


def connect_peers(seg, rank, world, cap_rows, all_gather_object, snapshot_bytes=0):
    """Sets up the exchange over NVLink peer memory (csrc/dsc_comm.cuh) for `seg` (a SegmentFunction): creates this rank's
    window, exchanges the CUDA IPC handles with `all_gather_object(obj) -> list of every rank's obj in rank order`
    (e.g. a wrapper of torch.distributed.all_gather_object) and maps the peers' windows.  Every rank must call it.
    snapshot_bytes > 0 adds the region of the sharded snapshot commit (SegmentFunction.commit_snapshot_sharded)."""
    handle = seg.comm_create(rank, world, cap_rows, snapshot_bytes) if snapshot_bytes else seg.comm_create(rank, world, cap_rows)
    handles = all_gather_object(handle)
    if len(handles) != world or any(len(h) != 64 for h in handles):
        raise ValueError("all_gather_object must return one 64-byte handle per rank")
    if handles[rank] != handle:
        raise ValueError("all_gather_object returned the handles in the wrong order")
    if world > 1:
        seg.comm_connect(handles)
    return handles
