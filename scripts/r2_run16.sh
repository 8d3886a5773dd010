#!/bin/bash
timeout 900 python -m pytest tests/test_staged_gpu.py tests/test_bench_configs_gpu.py tests/test_comm_gpu.py -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; tail -3 gpurun_out/r2q_tests.log
timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches.csv python scripts/profile_target.py C4 3 hot > /dev/null 2>&1
grep "tet_staged\|tet_exact\|finalize" gpurun_out/r2q_launches.csv | awk -F'","' '{print $5, $NF}' | tail -3
PART=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_comm_launches_part8.csv python scripts/profile_comm.py C4 3 > /dev/null 2>&1
grep "comm_\|force_geom\|tet_staged\|tet_exact" gpurun_out/r2q_comm_launches_part8.csv | awk -F'","' '{print $5, $NF}' | tail -5
