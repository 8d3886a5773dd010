#!/bin/bash
P=3d-image-segmentation-with-mesh_b200
run() { echo "== lib=$1"; DSCGPU_LIBRARY=$1 timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -3; }
run ""
run $P/libdscgpu_p8.so
run $P/libdscgpu_p4.so
