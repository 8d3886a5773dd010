#!/bin/bash
# One GPU pass of the round: the full GPU test suite, the default bench line, and the time line of one end-to-end step.
#   scripts/gpurun_retry.sh gpurun_out/check.log 2400 'bash scripts/r2_gpu_check.sh'
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/check_tests.log 2>&1; tail -3 gpurun_out/check_tests.log
timeout 600 python bench.py > gpurun_out/check_bench_c4_n1.json 2> gpurun_out/check_bench_c4_n1.err; tail -c 200 gpurun_out/check_bench_c4_n1.json; tail -2 gpurun_out/check_bench_c4_n1.err
DSCGPU_E2E_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/check_trace.err
L=$(grep -n "start (copy" gpurun_out/check_trace.err | tail -2 | head -1 | cut -d: -f1); sed -n "$((L-1)),$((L+18))p" gpurun_out/check_trace.err > gpurun_out/check_e2e_timeline.txt; cat gpurun_out/check_e2e_timeline.txt
