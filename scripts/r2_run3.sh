#!/bin/bash
# staged kernel: parity tests, timing of library builds, launch list with per-kernel durations, ncu of the default build
timeout 600 python -m pytest tests/test_staged_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; tail -3 gpurun_out/r2c_tests.log
P=3d-image-segmentation-with-mesh_b200
for lib in "" $P/libdscgpu_p8.so; do
    echo "== lib=$lib"
    DSCGPU_LIBRARY=$lib DSCGPU_TET_VARIANT=147584 timeout 200 python scripts/profile_target.py C4 6 hot 2>&1 | grep "tet ms" | cut -c1-30
    DSCGPU_LIBRARY=$lib DSCGPU_TET_VARIANT=147584 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches.csv python scripts/profile_target.py C4 3 hot > /dev/null 2>&1
    grep "tet_staged\|tet_exact" gpurun_out/r2c_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tet_staged -s 1 -c 1 -f -o gpurun_out/prof_tet_staged_r2c env DSCGPU_LIBRARY=$P/libdscgpu_p8.so DSCGPU_TET_VARIANT=147584 python scripts/profile_target.py C4 3 hot > gpurun_out/r2c_ncu_full.log 2>&1; tail -2 gpurun_out/r2c_ncu_full.log
