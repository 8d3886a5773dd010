#!/bin/bash
timeout 600 python -m pytest tests/test_staged_gpu.py tests/test_comm_gpu.py tests/test_incremental_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; tail -3 gpurun_out/r2i_tests.log
for part in 1 8; do
  echo "== PART=$part"
  PART=$part timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -3
  PART=$part timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_launches_p$part.csv python scripts/profile_target.py C4 3 hot > /dev/null 2>&1
  grep "tet_staged\|tet_exact\|finalize" gpurun_out/r2i_launches_p$part.csv | awk -F'","' '{print $5, $NF}' | tail -3
done
