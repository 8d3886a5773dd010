#!/bin/bash
# C5 (1024^3 u16, 16.1 M tets) on 8 GPUs, then C4 on 4
N=8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C5 --steps 10 --warmup 3 --no-parity > gpurun_out/r2k_bench_c5_n$N.json 2> gpurun_out/r2k_bench_c5_n$N.err; tail -c 400 gpurun_out/r2k_bench_c5_n$N.json; tail -3 gpurun_out/r2k_bench_c5_n$N.err
N=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-parity > gpurun_out/r2k_bench_c4_n$N.json 2> gpurun_out/r2k_bench_c4_n$N.err; tail -c 300 gpurun_out/r2k_bench_c4_n$N.json
