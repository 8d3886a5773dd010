"""CPU restatement (numpy) of the FP32 filter of the default tet kernel (csrc/dsc_kernels.cuh, tet_integrate_w2_kernel) against
the reference's FP64 sample position (DEMO/tet_dis_coord.hpp:27, truncation DEMO/image3d.cpp:366).

Phase A forms, per tet: O = floor(min corner), Q0 - 1/2 and E_k = (P_k - P_0)/(4n) rounded to float, thr = 1/2 - delta with
delta = (48 Rmax + 8) 2^-24.  Phase B evaluates q = Q0 - 1/2 + eps O + a1 E1 + a2 E2 + a3 E3 with four FMAs, rounds with the
1.5*2^23 trick and accepts the voxel when |q - round(q)| < thr in all three coordinates.  Checked here, on random tets
including ones with nodes exactly on the faces of the volume:
  * the error bound the guard band rests on: |q - (c - O - 1/2)| <= (41 Rmax + 5) 2^-24, c the reference's FP64 coordinate;
  * the claim itself: every accepted sample has O + round(q) == trunc(c) and lies inside the volume;
  * the fraction of samples sent to the exact path stays small (the band is not vacuous).
float32 FMA is emulated as float32(float64(a) * float64(b) + float64(c)): the product is exact, the sum is rounded twice, which
can differ from a true FMA by one float ulp in rare ties -- far inside the slack between the two bounds."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32
MAGIC = F(12582912.0)  # 1.5 * 2^23


def _table():
    txt = open(os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200/csrc/dsc_tables.h")).read()
    body = re.sub(r"//[^\n]*", "", re.search(r"DSC_TET_BARY\[[^\]]*\] = \{([^}]*)\}", txt).group(1))
    b = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()], dtype=np.float64).reshape(-1, 4)
    offs = [(L * (L + 1) // 2) ** 2 for L in range(10)]
    return b, offs


def _fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _tet_level(v):
    # upper_bound(dis_coord_size = n^3, v) - 1 clamped at 0 (DEMO/image3d.cpp:355-358): n = floor(cbrt(v)) in [1, 9]
    n = 1
    while n < 9 and (n + 1) ** 3 <= v:
        n += 1
    return n - 1


def _volume(p):
    a, b, c = p[0] - p[3], p[1] - p[3], p[2] - p[3]
    return abs(np.dot(a, np.cross(b, c))) / 6.0


def _axis_in(lo, hi, dim):
    return lo >= -1.0e-7 and hi <= dim + 1.0e-7 and hi - lo >= 1.0e-4 * dim


def _run(tets_xyz, dims, b, offs):
    n_acc = n_all = 0
    worst = 0.0
    for p in tets_xyz:
        v = _volume(p)
        L = _tet_level(v)
        if L >= 5:
            continue
        n = L + 1
        lo, hi = p.min(axis=0), p.max(axis=0)
        strict = all(lo[k] >= 1.5 and hi[k] < dims[k] - 1.5 for k in range(3))
        if not (strict or all(_axis_in(lo[k], hi[k], dims[k]) for k in range(3))):
            continue  # exact path in the kernel
        O = np.floor(lo)
        rmax = float((hi - O).max())
        if rmax >= 100.0:
            continue
        rows = b[offs[L]:offs[L + 1]]
        a = np.rint(4.0 * n * rows)
        eps = (((rows[:, 0] + rows[:, 1]) + rows[:, 2]) + rows[:, 3] - 1.0).astype(F)
        a1, a2, a3 = a[:, 1].astype(F), a[:, 2].astype(F), a[:, 3].astype(F)
        inv4n = 1.0 / (4.0 * n)
        thr = F(0.5) - F(F(F(48.0) * F(rmax)) + F(8.0)) * F(5.9604644775390625e-08)
        bound = (41.0 * rmax + 5.0) * 2.0 ** -24
        ok = np.ones(len(rows), dtype=bool)
        idx = np.zeros((len(rows), 3))
        cs = np.zeros((len(rows), 3))
        for k in range(3):
            c = ((p[0, k] * rows[:, 0] + p[1, k] * rows[:, 1]) + p[2, k] * rows[:, 2]) + p[3, k] * rows[:, 3]
            q0 = F((p[0, k] - O[k]) - 0.5)
            e1, e2, e3 = F((p[1, k] - p[0, k]) * inv4n), F((p[2, k] - p[0, k]) * inv4n), F((p[3, k] - p[0, k]) * inv4n)
            ones = np.ones(len(rows), dtype=F)
            q = _fma32(a3, e3 * ones, _fma32(a2, e2 * ones, _fma32(a1, e1 * ones, _fma32(eps, F(O[k]) * ones, q0 * ones))))
            worst = max(worst, float(np.abs(q.astype(np.float64) - (c - O[k] - 0.5)).max() / bound))
            t = (q + MAGIC).astype(F)
            r = (t - MAGIC).astype(F)
            e = (q - r).astype(F)
            ok &= np.abs(e) < thr
            idx[:, k] = r.astype(np.float64) + O[k]
            cs[:, k] = c
        acc = ok
        assert np.array_equal(idx[acc], np.trunc(cs[acc])), "an accepted sample landed in another voxel than the reference's"
        assert np.all(idx[acc] >= 0) and np.all(idx[acc] <= np.array(dims) - 1)
        n_acc += int(acc.sum())
        n_all += len(rows)
    return n_acc, n_all, worst


def test_filter_bound_and_classification_interior_and_face_tets():
    b, offs = _table()
    rng = np.random.default_rng(11)
    dims = (64, 48, 80)
    tets = []
    for _ in range(1500):  # interior, arbitrary position
        c = rng.uniform(8, np.array(dims) - 8)
        tets.append(c + rng.uniform(-1, 1, (4, 3)) * rng.uniform(1.0, 6.0))
    for _ in range(1500):  # corners snapped onto the faces of the volume
        c = rng.uniform(0, np.array(dims))
        p = np.clip(c + rng.uniform(-1, 1, (4, 3)) * rng.uniform(1.0, 7.0), 0.0, np.array(dims, dtype=np.float64))
        tets.append(p)
    for _ in range(300):  # first layer of a grid mesh: exact lattice coordinates k * edge, a face plane at 0.0 and at dim
        edge = 7.641791044776119  # C4's edge: not a multiple of 1/(4n), or every sample would sit exactly on a voxel boundary
        i = np.array([rng.integers(0, int(d / edge) + 1) for d in dims])
        cell = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]]) if rng.random() < 0.5 else np.array([[0, 0, 0], [0, 1, 0], [0, 1, 1], [1, 1, 1]])
        p = np.minimum((i + cell) * edge, np.array(dims, dtype=np.float64))
        tets.append(p.astype(np.float64))
    n_acc, n_all, worst = _run(tets, dims, b, offs)
    assert n_all > 20000
    assert worst <= 1.0, worst                   # the proven bound holds on every sample
    assert n_acc / n_all > 0.97, n_acc / n_all   # and the band sends only a sliver of the samples to the exact path
    n_acc_i, n_all_i, _ = _run(tets[:1500], dims, b, offs)  # (tets clipped to the faces put more samples near voxel boundaries)
    assert n_acc_i / n_all_i > 0.999
