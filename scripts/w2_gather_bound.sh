#!/bin/bash
# Upper bound of what a perfect memory system would give the w2 tet kernel: builds the library twice with the kernel's gathers
# replaced by (1) loads from a 16 KB window of the volume that stays in L1, (2) loads from shared memory, and runs bench.py on
# each.  The results of those runs are INVALID (the voxels read are the wrong ones); only kernel_ms is of interest.
# Round 1, C4, B200: real gathers 0.374 ms, L1-resident 0.324 ms, shared memory 0.313 ms.
# Run from the repository root on a GPU box:  bash scripts/w2_gather_bound.sh
set -e
P=3d-image-segmentation-with-mesh_b200
cp $P/libdscgpu.so /tmp/libdscgpu.real.so
for f in 1 2; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -DW2_FAKE=$f -shared -Xcompiler -fPIC -Xcompiler -O2 \
        -ccbin /usr/bin/g++ -w -o $P/libdscgpu.so $P/csrc/dscgpu.cu
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('W2_FAKE=$f', d['kernel_ms'])"
done
cp /tmp/libdscgpu.real.so $P/libdscgpu.so
