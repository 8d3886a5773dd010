#!/bin/bash
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-parity > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; tail -c 700 gpurun_out/r2g_bench_n$N.json; tail -5 gpurun_out/r2g_bench_n$N.err
