#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_tests.log 2>&1; tail -3 gpurun_out/r2ad_tests.log
timeout 600 python bench.py > gpurun_out/r2ad_bench_c4_n1.json 2> gpurun_out/r2ad_bench_c4_n1.err; tail -c 200 gpurun_out/r2ad_bench_c4_n1.json; tail -2 gpurun_out/r2ad_bench_c4_n1.err
DSCGPU_E2E_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2ad_trace.err
L=$(grep -n "start (copy" gpurun_out/r2ad_trace.err | tail -2 | head -1 | cut -d: -f1); sed -n "$((L-1)),$((L+18))p" gpurun_out/r2ad_trace.err > gpurun_out/r2ad_e2e_timeline.txt; cat gpurun_out/r2ad_e2e_timeline.txt
