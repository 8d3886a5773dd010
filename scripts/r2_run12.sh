#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; tail -5 gpurun_out/r2l_tests.log
timeout 600 python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; tail -c 300 gpurun_out/r2l_bench_n1.json; tail -3 gpurun_out/r2l_bench_n1.err
