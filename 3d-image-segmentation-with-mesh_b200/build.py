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
DEPS = ["dscgpu.cu", "dsc_kernels.cuh", "dsc_tet_staged.cuh", "dsc_comm.cuh", "dsc_device.cuh", "dsc_tables.h", os.path.join("..", "..", "include", "dscgpu.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found: cannot build the CUDA library (there is no CPU fallback)")


def source_hash():
    import hashlib
    h = hashlib.sha256()
    for d in DEPS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(f.read())
    h.update(os.environ.get("W2_MIN_BLOCKS", "").encode())
    return h.hexdigest()


def needs_build():
    """Stale = the sources differ from the ones the library was built from (content hash in a sidecar file; modification
    times do not survive every way a tree gets copied to a GPU box)."""
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + ".srchash") as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build_variant(name, defines):
    """Development: an experimental copy of the library with extra -D flags, as lib<name>.so next to the real one (DSCGPU_LIBRARY)."""
    out = os.path.join(HERE, "lib%s.so" % name)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"] + ["-D" + d for d in defines] + [
        "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


def build(force=False, verbose=False):
    """Builds the library if it is missing or older than its sources.  Safe under concurrency (one process per GPU all import
    the package at once): an exclusive file lock serialises the check, and the compiler writes to a temporary file that is
    renamed into place, so no process ever maps a half-written library."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while this one waited
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17"] + (["-DW2_MIN_BLOCKS=" + os.environ["W2_MIN_BLOCKS"]] if os.environ.get("W2_MIN_BLOCKS") else []) + ["-shared",
                   "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++", "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed")
            if verbose:
                sys.stderr.write(r.stdout + r.stderr)
            os.replace(tmp, LIB)
            with open(LIB + ".srchash", "w") as f:
                f.write(source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
