#!/bin/bash
# A/B of experimental builds of the library (build.py build_variant -> lib<name>.so): staged-kernel parity tests with the default library, then
# the tet pass timing of the default and of each variant, twice, interleaved:  bash scripts/ab_library.sh name1 name2 ...
timeout 900 python -m pytest tests/test_staged_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for v in "" "$@"; do
  lib=""; [ -n "$v" ] && lib=$PWD/3d-image-segmentation-with-mesh_b200/lib$v.so
  echo "library ${v:-default}: $(DSCGPU_LIBRARY=$lib timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep 'tet ms' | cut -c8-14 | tail -3 | tr '\n' ' ')"
done; done
