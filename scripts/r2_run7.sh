#!/bin/bash
# multi-GPU: exchange-kernel tests on one GPU, then bench at N=2 (and N=1 for reference)
timeout 600 python -m pytest tests/test_comm_gpu.py tests/test_incremental_gpu.py -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; tail -5 gpurun_out/r2g_tests.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; tail -c 600 gpurun_out/r2g_bench_n$N.json; tail -5 gpurun_out/r2g_bench_n$N.err
