#!/usr/bin/env python3
"""Structure of the reference's tetrahedron sample tables (DEMO/tet_dis_coord.cpp), as used by the staged tet kernel.

Checks, for levels n = 1..5, that every row (b0,b1,b2,b3) is a_k/(4n) + eta_k with integers a = (4n-a1-a2-a3, 4j+oy, 4i+ox, 4k+oz),
(ox,oy,oz) one of six offset vectors, and prints per level the (cell, offset) slots, eta / eps classes and whether level n's slots
are a subset of level n+1's slots (cell index and offset id).
"""
import importlib.util, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("gt", os.path.join(HERE, "..", "tools", "gen_tables.py"))
gt = importlib.util.module_from_spec(spec); spec.loader.exec_module(gt)
OFFS = [(1, 1, 1), (2, 1, 2), (1, 2, 3), (3, 2, 1), (2, 3, 2), (3, 3, 3)]  # (ox, oy, oz)
slots = {}
for n in range(1, 6):
    rows = gt.tet_level(n)
    S = []
    eps_vals = {}
    for r in rows:
        a = [int(round(x * 4 * n)) for x in r]
        assert sum(a) == 4 * n and min(a) >= 1, (n, r, a)
        eta = [x - ak / (4.0 * n) for x, ak in zip(r, a)]
        assert max(abs(e) for e in eta) <= 5.0e-7
        a1, a2, a3 = a[1], a[2], a[3]   # y, x, z
        o = (a2 % 4 or 4, a1 % 4 or 4, a3 % 4 or 4)
        # offsets are 1..3 so a % 4 is never 0
        assert o in OFFS, (n, a, o)
        i, j, k = (a2 - o[0]) // 4, (a1 - o[1]) // 4, (a3 - o[2]) // 4
        S.append(((i, j, k), OFFS.index(o)))
        eps = sum(r) - 1.0
        key = round(eps * 3e6)
        eps_vals[key] = eps_vals.get(key, 0) + 1
    slots[n] = S
    print("n=%d rows=%d cells=%d eps classes (x 1/3 e-6): %s" % (n, len(rows), len(set(c for c, _ in S)), dict(sorted(eps_vals.items()))))
    assert len(set(S)) == len(S)
for n in range(1, 5):
    print("n=%d subset of n=%d: %s" % (n, n + 1, set(slots[n]) <= set(slots[n + 1])))
# row order of level 4 as (cell, offset)
print(slots[3])
