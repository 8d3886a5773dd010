#!/bin/bash
P=3d-image-segmentation-with-mesh_b200
run() { echo "== lib=$1 kb=$2"; DSCGPU_LIBRARY=$1 DSCGPU_STG_SLOT_KB=$2 DSCGPU_TET_VARIANT=147584 timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -3; }
run "" 24
run $P/libdscgpu_p8.so 24
run $P/libdscgpu_p4.so 24
