#!/bin/bash
# full GPU tests; bench lines for C4 (default), C2, C3 and the z-ray mean mode
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; tail -5 gpurun_out/r2o_tests.log
timeout 600 python bench.py > gpurun_out/r2o_bench_c4_n1.json 2> gpurun_out/r2o_bench_c4_n1.err; tail -c 200 gpurun_out/r2o_bench_c4_n1.json; tail -2 gpurun_out/r2o_bench_c4_n1.err
for c in C2 C3; do timeout 600 python bench.py --config $c > gpurun_out/r2o_bench_${c}_n1.json 2> gpurun_out/r2o_bench_${c}_n1.err; tail -c 200 gpurun_out/r2o_bench_${c}_n1.json; done
timeout 600 python bench.py --mean-mode zray --no-cpu-baseline > gpurun_out/r2o_bench_c4_n1_zray.json 2> gpurun_out/r2o_bench_c4_n1_zray.err; tail -c 200 gpurun_out/r2o_bench_c4_n1_zray.json; tail -2 gpurun_out/r2o_bench_c4_n1_zray.err
