#!/bin/bash
timeout 1500 python -m pytest tests/test_dropin_gpu.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; tail -5 gpurun_out/r2m_tests.log
