#!/bin/bash
# Round-2 measurement: GPU tests, default and staged bench lines, ncu full capture of the staged tet kernel.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; tail -3 gpurun_out/r2b_tests.log
bash scripts/ab_variants.sh 16512 147584
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tet_staged -s 1 -c 1 -f -o gpurun_out/prof_tet_staged_r2b env DSCGPU_TET_VARIANT=147584 python scripts/profile_target.py C4 3 hot > gpurun_out/r2b_ncu_full.log 2>&1; tail -2 gpurun_out/r2b_ncu_full.log
