#!/bin/bash
# full GPU tests (incl. the drop-in run against the live reference objects) and the default bench line
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; tail -5 gpurun_out/r2w_tests.log
grep -h "internal forces\|TOTAL\|refresh" gpurun_out/dropin_edge3.log 2>/dev/null | tail -8
timeout 600 python bench.py > gpurun_out/r2w_bench_c4_n1.json 2> gpurun_out/r2w_bench_c4_n1.err; tail -c 300 gpurun_out/r2w_bench_c4_n1.json; tail -2 gpurun_out/r2w_bench_c4_n1.err
