#!/bin/bash
# A/B of the producer / consumer warp split of the staged kernel (slack slots in the ring)
for v in "" _p9c7 _p10c6 _p8c7; do
  echo "variant libdscgpu$v"
  DSCGPU_LIBRARY=$PWD/3d-image-segmentation-with-mesh_b200/libdscgpu$v.so timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -3
done
