"""The C-ABI library loads without a GPU and exports every symbol include/dscgpu.h declares;
compute calls fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dscgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dscgpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for need in ("dscgpu_ctx_create", "dscgpu_set_snapshot", "dscgpu_tet_integrate", "dscgpu_zray_means", "dscgpu_zray_energy",
                 "dscgpu_face_forces", "dscgpu_image_force_eval", "dscgpu_build_sum_table", "dscgpu_points_in_tets"):
        assert need in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device(pkg):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.DscGpuError):
        pkg.Image3D(np.zeros((4, 4, 4), np.float32))


def test_product_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, "3d-image-segmentation-with-mesh_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f)).read()
                assert "flat_oracle" not in src and "libflatoracle" not in src and "from oracle" not in src and "import oracle" not in src, f
