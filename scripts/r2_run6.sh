#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 1500 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err
