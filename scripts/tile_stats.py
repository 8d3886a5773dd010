#!/usr/bin/env python3
"""Voxel-box statistics of the 32-tet tiles of a BASELINE config mesh (host-side planning aid for the staged tet kernel).

For each tile of 32 consecutive tets: the box of voxels its samples can touch (sample AABB: every barycentric weight of
a level-n row is >= 1/(4n), so a sample stays (hi-lo)/(4n) away from the faces of the tet's AABB), x origin aligned down
to 16 bytes.  Prints the distribution of box sizes, the level mix and how many tiles fit a given slot size.
usage: tile_stats.py [C4] [tile=32]
"""
import importlib.util, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("syn", os.path.join(HERE, "..", "3d-image-segmentation-with-mesh_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = syn.CONFIGS[name]
n = cfg["n"]; esz = np.dtype(cfg["dtype"]).itemsize
edge = syn.edge_for_cells(n, cfg["cells"])
pos, tets, grid = syn.grid_mesh((n, n, n), edge, 0.2, 1234)
base = syn.boundary_layer_labels(tets, grid)
P = pos[tets.astype(np.int64)]                       # [m,4,3]
a, b, c, d = P[:, 0], P[:, 1], P[:, 2], P[:, 3]
vol = np.abs(np.einsum("ij,ij->i", a - d, np.cross(b - d, c - d))) / 6.0
lev = np.clip(np.floor(np.cbrt(vol) + 1e-12).astype(int), 1, 9)
lo, hi = P.min(axis=1), P.max(axis=1)
shrink = (hi - lo) / (4.0 * lev[:, None]) * 0.999
slo, shi = np.floor(lo + shrink).astype(int), np.floor(hi - shrink).astype(int)
act = base != 0
m = len(tets); nt = (m + T - 1) // T
pad = nt * T - m
def tile_red(x, fn, fill):
    x = np.concatenate([x, np.full((pad,) + x.shape[1:], fill, x.dtype)]).reshape(nt, T, *x.shape[1:])
    return fn(x, axis=1)
big = 10**9
tlo = tile_red(np.where(act[:, None], slo, big), np.min, big)
thi = tile_red(np.where(act[:, None], shi, -big), np.max, -big)
nact = tile_red(act.astype(int), np.sum, 0)
al = 16 // esz
tlo[:, 0] = (tlo[:, 0] // al) * al
ext = thi - tlo + 1
ok = nact > 0
ext = ext[ok]
ex = ((ext[:, 0] + al - 1) // al) * al
byt = ex * ext[:, 1] * ext[:, 2] * esz
print(name, "tets", m, "active", act.sum(), "tiles", nt, "tiles with work", ok.sum(), "elem bytes", esz)
print("levels of active tets:", {int(k): int(v) for k, v in zip(*np.unique(lev[act], return_counts=True))})
for q in (50, 90, 95, 99, 100):
    print("pct %3d: ex %4d ey %3d ez %3d bytes %8d" % (q, np.percentile(ex, q), np.percentile(ext[:, 1], q), np.percentile(ext[:, 2], q), np.percentile(byt, q)))
for slot in (20, 22, 24, 25, 26, 27, 28, 32, 36):
    print("slot %2d KB: %5.1f %% of tiles fit; mean box of those %.1f KB" % (slot, 100.0 * (byt <= slot * 1024).mean(), byt[byt <= slot * 1024].mean() / 1024))
print("sum of boxes that fit 27 KB: %.3f GB" % (byt[byt <= 27 * 1024].sum() / 1e9))
n3 = (lev ** 3) * act
print("samples", n3.sum())
# lane efficiency of per-level passes inside a tile
levt = np.concatenate([np.where(act, lev, 0), np.zeros(pad, int)]).reshape(nt, T)
steps = sum(((levt == L).any(axis=1)) * L ** 3 for L in range(1, 10))
print("thread-per-tet steps per tile (sum over levels present): mean %.1f; lane efficiency %.3f" % (steps[ok].mean(), n3.sum() / (steps.sum() * T)))
