#!/bin/bash
# staged-kernel parity tests + tet pass timing of the default library
timeout 900 python -m pytest tests/test_staged_gpu.py tests/test_bench_configs_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python scripts/profile_target.py C4 5 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -4
