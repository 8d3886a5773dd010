"""Builds libdscgpu.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

The library is compiled with -fmad=false: positions, volumes, areas and barycentric tests feed
integer decisions that must match the reference's non-contracted x86-64 arithmetic bit for bit
(see csrc/dsc_device.cuh).  -lineinfo keeps ncu's source page usable.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdscgpu.so")
SOURCES = ["dscgpu.cu"]
DEPS = ["dscgpu.cu", "dsc_kernels.cuh", "dsc_device.cuh", "dsc_tables.h", os.path.join("..", "..", "include", "dscgpu.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found: cannot build the CUDA library (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"] + (["-DW2_MIN_BLOCKS=" + os.environ["W2_MIN_BLOCKS"]] if os.environ.get("W2_MIN_BLOCKS") else []) + ["-shared",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
