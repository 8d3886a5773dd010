"""Sample tables: regenerated == reference literals (container) and pinned by checksum (everywhere)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHA = "344a97ea2ea4c08e958805b52c22915cbcea9c7bf9d06fe467d5a860a05605d0"


def _header_sha(path):
    m = re.search(r"sha256 of the packed doubles: ([0-9a-f]{64})", open(path).read())
    return m.group(1)


def test_generated_headers_are_pinned():
    for p in ("oracle/tables_gen.h", "3d-image-segmentation-with-mesh_b200/csrc/dsc_tables.h"):
        assert _header_sha(os.path.join(ROOT, p)) == SHA


def test_generator_reproduces_checksum():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py")], capture_output=True, text=True, check=True).stdout
    assert SHA in out


@pytest.mark.skipif(not os.path.exists("/root/reference/DEMO/tet_dis_coord.cpp"), reason="reference tree only exists in the build container")
def test_tables_equal_reference_literals():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tables.py"), "--check", "/root/reference"], capture_output=True,
                         text=True, check=True).stdout
    assert "tables identical to the reference literals: 2025 tet points, 1015 tri points, 42 gauss points" in out


def test_level_offsets_closed_form():
    # the kernels use offset(L) = (L(L+1)/2)^2 and n3 = (L+1)^3 instead of a table
    txt = open(os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200/csrc/dsc_tables.h")).read()
    offs = [int(x) for x in re.search(r"DSC_TET_LEVEL_OFFSET\[[^\]]*\] = \{([^}]*)\}", txt).group(1).split(",")]
    assert offs == [(L * (L + 1) // 2) ** 2 for L in range(10)]


def _tet_table():
    import numpy as np
    txt = open(os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200/csrc/dsc_tables.h")).read()
    body = re.search(r"DSC_TET_BARY\[[^\]]*\] = \{([^}]*)\}", txt).group(1)
    body = re.sub(r"//[^\n]*", "", body)  # "// n = k" markers between the levels
    b = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()], dtype=np.float64).reshape(-1, 4)
    offs = [int(x) for x in re.search(r"DSC_TET_LEVEL_OFFSET\[[^\]]*\] = \{([^}]*)\}", txt).group(1).split(",")]
    return b, offs


def test_lattice_form_and_minimum_weight():
    """What the w2 kernel's FP32 filter and its in-volume argument assume about the reference's 6-decimal sample table
    (csrc/dsc_kernels.cuh): b_k = a_k/(4n) + eta_k with integers a_k >= 1 summing to 4n, |eta_k| <= 5e-7,
    |sum b_k - 1| <= 2.1e-6.  The library checks the same when it builds the table (dscgpu_create)."""
    import numpy as np
    b, offs = _tet_table()
    assert len(b) == 2025
    for L in range(9):
        n = L + 1
        rows = b[offs[L]:offs[L + 1]]
        assert len(rows) == n ** 3
        a = np.rint(4.0 * n * rows)
        assert a.min() >= 1 and np.all(a.sum(axis=1) == 4 * n)
        assert np.abs(rows - a / (4.0 * n)).max() <= 5.0e-7
        assert np.abs(((rows[:, 0] + rows[:, 1]) + rows[:, 2]) + rows[:, 3] - 1.0).max() <= 2.1e-6


def test_in_volume_argument_for_tets_on_the_faces():
    """Tets with -1e-7 <= lo, hi <= dim + 1e-7 and hi - lo >= 1e-4 dim on an axis: every sample coordinate, formed in the
    reference's order (DEMO/tet_dis_coord.hpp:27), lies strictly inside (0, dim) -- so truncation equals floor and no bounds
    test is needed on the filtered path.  Adversarial corners: three on the face and one at the minimum extent, both faces,
    plus random ones; dims up to 2048."""
    import numpy as np
    b, offs = _tet_table()
    rng = np.random.default_rng(5)
    for dim in (16, 100, 512, 1024, 2048):
        ext = 1.0e-4 * dim
        cases = [np.array([dim + 1e-7] * 3 + [dim + 1e-7 - ext]), np.array([dim + 1e-7 - ext] * 3 + [dim + 1e-7]),
                 np.array([-1e-7] * 3 + [-1e-7 + ext]), np.array([-1e-7 + ext] * 3 + [-1e-7]),
                 np.array([-1e-7, dim + 1e-7, 0.0, float(dim)])]
        for _ in range(200):
            lo = rng.uniform(-1e-7, dim - ext)
            hi = rng.uniform(lo + ext, dim + 1e-7)
            c = rng.uniform(lo, hi, 4)
            c[rng.integers(0, 4)] = lo
            c[(np.argmin(c) + 1 + rng.integers(0, 3)) % 4] = hi
            cases.append(c)
        for L in range(5):  # the filtered path serves n <= 5
            rows = b[offs[L]:offs[L + 1]]
            for c in cases:
                for perm in (c, c[::-1], np.roll(c, 1)):
                    x = ((perm[0] * rows[:, 0] + perm[1] * rows[:, 1]) + perm[2] * rows[:, 2]) + perm[3] * rows[:, 3]
                    assert x.min() > 0.0 and x.max() < dim, (dim, L, perm)
                    assert np.array_equal(np.trunc(x), np.floor(x)) and np.trunc(x).max() <= dim - 1
