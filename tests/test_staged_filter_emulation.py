"""CPU restatement (numpy float32) of the lattice arithmetic of the staged tet kernel (csrc/dsc_tet_staged.cuh, stg_block) against
the reference's FP64 sample position (DEMO/tet_dis_coord.hpp:27, truncation DEMO/image3d.cpp:366).

The kernel never reads the sample table on its fast path: it rebuilds every row from the structure of the reference's
generator (DEMO/image3d.cpp:256-306) -- cell (i,j,k), one of six centroid offsets -- with
    q = (P0 - O - 1/2) + j D1 + i D2 + k D3 + (oy E1 + ox E2 + oz E3) [+/- 1e-6 P0 on eight rows of level 3].
Checked here:
  * the slot enumeration of stg_block (loops j, i, k, offsets kept per cell) is a permutation of the table's rows, for the
    level of the instruction stream and for the level below riding it (stg_need);
  * the error bound behind stg_threshold: |q_float - (c - O - 1/2)| <= delta, c the reference's FP64 coordinate;
  * the claim itself: in every tet whose largest |q - round(q)| stays below 1/2 - delta all voxels equal trunc(c), and stay
    inside the tet's voxel box;
  * the guard band is not vacuous (few tets flagged).
float32 FMA is emulated as float32(float64(a) * float64(b) + float64(c)) (product exact, sum rounded twice: can differ from a
true FMA by one float ulp in rare ties, far inside the slack)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32
MAGIC = F(12582912.0)
OFF = [(1, 1, 1), (2, 1, 2), (1, 2, 3), (3, 2, 1), (2, 3, 2), (3, 3, 3)]  # stg_off: (ox, oy, oz)


def _table():
    txt = open(os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200/csrc/dsc_tables.h")).read()
    body = re.sub(r"//[^\n]*", "", re.search(r"DSC_TET_BARY\[[^\]]*\] = \{([^}]*)\}", txt).group(1))
    b = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()], dtype=np.float64).reshape(-1, 4)
    offs = [(L * (L + 1) // 2) ** 2 for L in range(10)]
    return b, offs


def _need(s, t):
    return s + (1 if t == 0 else (2 if t <= 4 else 3))


def _eps3(i, j, k, t):
    a1, a2, a3 = 4 * j + OFF[t][1], 4 * i + OFF[t][0], 4 * k + OFF[t][2]
    a = [12 - a1 - a2 - a3, a1, a2, a3]
    c = sum((x % 3 == 2) - (x % 3 == 1) for x in a)
    assert c in (-3, 0, 3)
    return c // 3


def _slots(NB):
    """(i, j, k, t) in the order stg_block visits them."""
    out = []
    for j in range(NB):
        for i in range(NB - j):
            for k in range(NB - j - i):
                s = i + j + k
                m = 6 if NB - s >= 3 else (5 if NB - s == 2 else 1)
                for t in range(m):
                    out.append((i, j, k, t))
    return out


def _row_of_slot(n, slot):
    """barycentric numerators (a0, a1, a2, a3) of a slot at level n"""
    i, j, k, t = slot
    a1, a2, a3 = 4 * j + OFF[t][1], 4 * i + OFF[t][0], 4 * k + OFF[t][2]
    return (4 * n - a1 - a2 - a3, a1, a2, a3)


def test_slots_are_the_table_rows():
    b, offs = _table()
    for NB in (1, 2, 3, 4):
        for n in (NB, NB - 1):
            if n < 1:
                continue
            rows = b[offs[n - 1]:offs[n - 1] + n ** 3]
            want = sorted(tuple(int(round(x * 4 * n)) for x in r) for r in rows)
            got = sorted(_row_of_slot(n, s) for s in _slots(NB) if _need(s[0] + s[1] + s[2], s[3]) <= n)
            assert got == want, (NB, n)
        assert len(_slots(NB)) == NB ** 3
    # eps of the level-3 literals: -/+1e-6 on eight rows, 0 elsewhere, as stg_eps3 says
    rows3 = {tuple(int(round(x * 12)) for x in r): float(sum(r) - 1.0) for r in b[offs[2]:offs[2] + 27]}
    cnt = 0
    for s in _slots(3):
        e = rows3[_row_of_slot(3, s)]
        assert abs(e - 1.0e-6 * _eps3(*s)) < 1e-12
        cnt += _eps3(*s) != 0
    assert cnt == 8
    # levels 1, 2, 4: the literals are exact binary fractions
    for n in (1, 2, 4):
        rows = b[offs[n - 1]:offs[n - 1] + n ** 3]
        assert np.array_equal(rows * (4 * n), np.round(rows * (4 * n)))


def _fma32(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def _tet_level(v):
    n = 1
    while n < 9 and (n + 1) ** 3 <= v:
        n += 1
    return n


def _emulate(p, n, NB, b, offs):
    """Returns (largest |q - round(q)|, thr, voxels [n3,3] in table order, reference voxels, worst error / delta)."""
    lo, hi = p.min(axis=0), p.max(axis=0)
    org = np.floor(lo)
    rmax = float((hi - org).max())
    diam = float((hi - lo).max())
    delta = 5.9604644775390625e-08 * (6.0 * (rmax + 1.0) + 6.0 + 10.0 * diam) + (1.01e-6 * diam if n == 3 else 0.0) + 1.0e-9
    thr = F(0.5 - delta) * F(0.99999988079071044921875)
    inv = 1.0 / n
    q0 = ((p[0] - org) - 0.5).astype(F)
    d = [((p[m] - p[0]) * inv).astype(F) for m in (1, 2, 3)]
    e = [F(0.25) * x for x in d]
    off = []
    for o in OFF:
        fx, fy, fz = F(o[0]), F(o[1]), F(o[2])
        off.append(np.array([_fma32(fz, e[2][c], _fma32(fx, e[1][c], fy * e[0][c])) for c in range(3)], F))
    ep = (1.0e-6 * p[0]).astype(F) if n == 3 else np.zeros(3, F)
    rows = b[offs[n - 1]:offs[n - 1] + n ** 3]
    rowmap = {tuple(int(round(x * 4 * n)) for x in r): r for r in rows}
    emax = 0.0
    worst = 0.0
    vox, ref = [], []
    for j in range(NB):
        cj = np.array([_fma32(F(j), d[0][c], q0[c]) for c in range(3)], F) if j else q0
        for i in range(NB - j):
            ci = np.array([_fma32(F(i), d[1][c], cj[c]) for c in range(3)], F) if i else cj
            for k in range(NB - j - i):
                ck = np.array([_fma32(F(k), d[2][c], ci[c]) for c in range(3)], F) if k else ci
                s = i + j + k
                m = 6 if NB - s >= 3 else (5 if NB - s == 2 else 1)
                for t in range(m):
                    q = ck + off[t]
                    if NB >= 3 and (NB == 3 or _need(s, t) <= 3):
                        es = _eps3(i, j, k, t)
                        if es > 0:
                            q = q + ep
                        elif es < 0:
                            q = q - ep
                    tt = q + MAGIC
                    r = tt - MAGIC
                    err = q - r
                    valid = _need(s, t) <= n
                    emax = max(emax, float(np.abs(err).max()))  # the kernel looks at every slot, valid for this tet or not
                    if not valid:
                        continue
                    row = rowmap[_row_of_slot(n, (i, j, k, t))]
                    c = ((p[0] * row[0] + p[1] * row[1]) + p[2] * row[2]) + p[3] * row[3]
                    worst = max(worst, float(np.abs(q.astype(np.float64) - (c - org - 0.5)).max()) / delta)
                    vox.append(org + r.astype(np.float64))
                    ref.append(np.trunc(c))
    return emax, float(thr), np.array(vox), np.array(ref), worst, lo, hi


def test_filter_bound_and_voxels():
    b, offs = _table()
    rng = np.random.default_rng(20261020)
    dims = 512
    n_tets = 0
    flagged = 0
    worst = 0.0
    for it in range(1500):
        edge = rng.choice([2.2, 3.4, 4.3, 5.2, 7.64])
        c0 = rng.uniform(12.0, dims - 12.0, 3)
        p = c0 + rng.uniform(-0.75, 0.75, (4, 3)) * edge
        if it % 7 == 0:  # a node exactly on a face of the volume, like the first layer of a DSC mesh
            p = p - p.min(axis=0)[None, :] * np.array([1.0, 0.0, 0.0])
        a_, b_, c_ = p[0] - p[3], p[1] - p[3], p[2] - p[3]
        v = abs(np.dot(a_, np.cross(b_, c_))) / 6.0
        n = _tet_level(v)
        if n > 4 or v < 0.05:
            continue
        for NB in (n, n + 1):
            if NB > 4:
                continue
            emax, thr, vox, ref, w, lo, hi = _emulate(p, n, NB, b, offs)
            n_tets += 1
            worst = max(worst, w)
            assert w <= 1.0, (it, n, NB, w)
            if emax < thr:
                assert np.array_equal(vox, ref), (it, n, NB)
                f = 0.9 / (4 * n)
                blo = np.floor(lo + f * (hi - lo) - 1.0e-3)
                bhi = np.floor(hi - f * (hi - lo) + 1.0e-3)
                assert (vox >= blo).all() and (vox <= bhi).all(), (it, n, NB)
            else:
                flagged += 1
    assert n_tets > 1000
    assert flagged < 0.05 * n_tets, (flagged, n_tets)
    assert worst > 0.02  # the bound is within two orders of magnitude of what occurs
