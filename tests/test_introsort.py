"""The z-ray kernel sorts hits with a restatement of libstdc++'s std::sort (csrc/dsc_device.cuh);
ties (equal z, different b_in) must end up exactly where the real std::sort puts them.  The
same __host__ __device__ code is exported for the host and fuzzed against std::sort here."""
import numpy as np
import pytest


@pytest.mark.parametrize("n", [0, 1, 2, 3, 15, 16, 17, 18, 26, 33, 34, 64, 95, 96, 200, 1000, 5000])
def test_matches_std_sort(pkg, flat, n):
    rng = np.random.default_rng(100 + n)
    for trial in range(60 if n <= 200 else 6):
        span = [2, 3, 5, 17, 1000][trial % 5]  # few distinct z => many ties
        z = rng.integers(-3, span, size=n).astype(np.int32)
        if trial % 7 == 3:
            z = np.sort(z)
        if trial % 7 == 5:
            z = np.sort(z)[::-1].copy()
        b = rng.integers(0, 2, size=n).astype(np.uint8)
        z1, b1 = pkg.host_sort_hits(z, b)
        z2, b2 = flat.std_sort_hits(z, b)
        assert np.array_equal(z1, z2)
        assert np.array_equal(b1, b2), "tie order differs from std::sort for n=%d" % n


def test_adversarial_patterns(pkg, flat):
    # organ-pipe and median-of-3 killer-like inputs push introsort towards its heap-sort fallback
    for n in (40, 64, 96, 512):
        pats = [np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]), np.tile([1, 0], n // 2), np.zeros(n), np.arange(n) % 3,
                np.concatenate([np.arange(0, n, 2), np.arange(1, n, 2)])]
        for p in pats:
            z = p.astype(np.int32)
            b = (np.arange(len(z)) % 2).astype(np.uint8)
            z1, b1 = pkg.host_sort_hits(z, b)
            z2, b2 = flat.std_sort_hits(z, b)
            assert np.array_equal(z1, z2) and np.array_equal(b1, b2)
