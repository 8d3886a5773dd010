"""B200-native image-energy path of the DSC tetrahedral-mesh segmentation loop.

Scope: per-tet voxel integration, per-phase means (tet-sampled and the
reference's z-ray/sum-table estimator), interface-face Gauss-quadrature forces, the energy
read-out -- as sm_100a CUDA kernels behind the C ABI in include/dscgpu.h.  This package holds
the CUDA sources (csrc/), the in-tree build (build.py), the ctypes host mirror of the
reference's segment_function / image3d call surface (host.py) and synthetic phantom / grid-mesh
generators (synthetic.py).  There is no CPU fallback: importing works anywhere, computing
needs the compiled library and a CUDA device.

The directory name is not a Python identifier; load it with importlib (see
__graft_entry__.load_package) under the module name ``dsc_b200``.
"""
from . import build as _build  # noqa: F401
from .host import (DscGpuError, Image3D, SegmentFunction, FORCE_ATOMIC, FORCE_GATHER, MEAN_TET, MEAN_ZRAY, library_path, load_library,
                   host_sort_hits)  # noqa: F401
from . import synthetic  # noqa: F401
from . import distributed  # noqa: F401
from . import mesh_io  # noqa: F401
