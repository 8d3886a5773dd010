"""Host-side emulation (numpy, float32 with correctly rounded fused multiply-adds) of the index arithmetic the bricked tet
kernels run on the device (csrc/dsc_tet_g32.cuh, tet_integrate_w2_kernel<BRICK>): lattice position in packed FP32, floor by the
1.5*2^23 trick, floor(r/4) and floor(r/2) from one more FMA on the rounded value, and the folded magic-bit constants of
idx = base + rx + 28 hx + 4 ry + (Sy-8) hy + 8 rz + (Sz-32) hz.  For every sample outside the guard band the emulated index must
be the brick index of floor(reference FP64 position)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_tables  # noqa: E402

f32 = np.float32
PAD = 8


def fma32(a, b, c):
    # a*b is exact in float64 for float32 inputs; the sum rounds once more in float64 (2^-53: irrelevant next to 2^-24)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def bits(x):
    return x.view(np.uint32).astype(np.uint64)


def brick_index(x, y, z, Sy, Sz):
    xp, yp, zp = x + PAD, y + PAD, z + PAD
    return (zp >> 2) * Sz + (yp >> 1) * Sy + ((xp >> 2) << 5) + ((zp & 3) << 3) + ((yp & 1) << 2) + (xp & 3)


def test_magic_constants():
    assert np.array([12582912.0], f32).view(np.uint32)[0] == 0x4B400000
    assert np.array([25165824.0], f32).view(np.uint32)[0] == 0x4BC00000
    assert np.array([50331648.0], f32).view(np.uint32)[0] == 0x4C400000
    assert float(f32(0.99999988079071044921875)) == 1 - 2.0 ** -23
    assert float(f32(0.999999940395355224609375)) == 1 - 2.0 ** -24
    assert 50331648.0 - 12582912.0 == 37748736.0


def test_floor_of_r_over_brick_from_one_fma():
    r = np.arange(-8, 200, dtype=np.float32)
    M = f32(12582912.0)
    t = (r + M).astype(f32)
    hx = fma32(t, np.full_like(t, f32(1 - 2.0 ** -23)), np.full_like(t, f32(37748736.0)))
    hy = fma32(t, np.full_like(t, f32(1 - 2.0 ** -24)), np.full_like(t, M))
    assert np.array_equal(bits(hx).astype(np.int64) - 0x4C400000, np.floor(r / 4).astype(np.int64))
    assert np.array_equal(bits(hy).astype(np.int64) - 0x4BC00000, np.floor(r / 2).astype(np.int64))


def test_emulated_index_matches_reference_floor():
    tet_tab, _ = gen_tables.build()
    X, Y, Z = 200, 150, 170
    NBX, NBY = (X + 2 * PAD + 3) // 4, (Y + 2 * PAD + 1) // 2
    Sy, Sz = 32 * NBX, 32 * NBX * NBY
    Ky, Kz = (Sy - 8) & 0xffffffff, (Sz - 32) & 0xffffffff
    MB, MB2, MB4 = 0x4B400000, 0x4BC00000, 0x4C400000
    fold = (MB * 13 + MB4 * (28 + Kz) + MB2 * Ky) & 0xffffffff
    M, K4, C4, K2 = f32(12582912.0), f32(1 - 2.0 ** -23), f32(37748736.0), f32(1 - 2.0 ** -24)
    rng = np.random.default_rng(0)
    checked = unsafe_n = 0
    for _ in range(1200):
        L = int(rng.integers(0, 9))
        n = L + 1
        tb = np.array(tet_tab[L])
        c = rng.uniform([-4, -4, -4], [X + 4, Y + 4, Z + 4])
        P = c + rng.uniform(-1, 1, (4, 3)) * rng.uniform(1, 12)
        lo, hi = P.min(0), P.max(0)
        if not (np.all(lo >= -6) and hi[0] < X + 6 and hi[1] < Y + 6 and hi[2] < Z + 6):
            continue
        o = np.floor(lo).astype(int)
        oa = np.array([((o[0] + PAD) & ~3) - PAD, ((o[1] + PAD) & ~1) - PAD, ((o[2] + PAD) & ~3) - PAD])
        rmax = (hi - oa).max()
        q0 = ((P[0] - oa) - 0.5).astype(f32)
        e = [((P[k] - P[0]) * (1.0 / (4 * n))).astype(f32) for k in (1, 2, 3)]
        thr = f32(0.5) - (f32(48.0) * f32(rmax) + f32(8.0)) * f32(5.9604644775390625e-08)
        a = np.rint(tb * 4 * n)
        eps = ((((tb[:, 0] + tb[:, 1]) + tb[:, 2]) + tb[:, 3]) - 1.0).astype(f32)
        a1, a2, a3 = [a[:, k].astype(f32) for k in (1, 2, 3)]
        base = (int(brick_index(oa[0], oa[1], oa[2], Sy, Sz)) - fold) & 0xffffffff
        m = len(tb)
        qs = []
        for d in range(3):
            q = fma32(eps, np.full(m, f32(oa[d])), np.full(m, q0[d], f32))
            for ak, ek in zip((a1, a2, a3), e):
                q = fma32(ak, np.full(m, ek[d]), q)
            qs.append(q)
        t = [(q + M).astype(f32) for q in qs]
        err = [fma32((tt - M).astype(f32), np.full(m, f32(-1)), q) for tt, q in zip(t, qs)]
        hx = fma32(t[0], np.full(m, K4), np.full(m, C4))
        hy = fma32(t[1], np.full(m, K2), np.full(m, M))
        hz = fma32(t[2], np.full(m, K4), np.full(m, C4))
        idx = (bits(t[0]) + 28 * bits(hx) + 4 * bits(t[1]) + Ky * bits(hy) + 8 * bits(t[2]) + Kz * bits(hz) + base) & 0xffffffff
        unsafe = np.maximum(np.maximum(np.abs(err[0]), np.abs(err[1])), np.abs(err[2])) >= thr
        pr = ((P[0][None, :] * tb[:, 0:1] + P[1][None, :] * tb[:, 1:2]) + P[2][None, :] * tb[:, 2:3]) + P[3][None, :] * tb[:, 3:4]
        fl = np.floor(pr).astype(int)
        ref = brick_index(fl[:, 0], fl[:, 1], fl[:, 2], Sy, Sz).astype(np.uint64)
        ok = ~unsafe
        assert np.array_equal(idx[ok], ref[ok])
        checked += int(ok.sum())
        unsafe_n += int(unsafe.sum())
    assert checked > 100000
    assert unsafe_n < checked // 500  # the guard band catches a fraction of a percent of the samples
