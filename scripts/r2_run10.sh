#!/bin/bash
timeout 600 python -m pytest tests/test_comm_gpu.py -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; tail -5 gpurun_out/r2j_tests.log
