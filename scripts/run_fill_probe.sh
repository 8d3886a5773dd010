#!/bin/bash
# runs the shared-memory fill probe over its modes (see tma_fill_probe.cu); output -> gpurun_out/fill_probe.log
mkdir -p gpurun_out
B=scripts/tma_fill_probe.bin
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem --format=csv
for mode in 0 1 2 3; do
  for work in 0 60 120; do
    timeout 60 $B $mode $work 8
  done
done
for mode in 0 2; do
  timeout 60 $B $mode 120 8 52 10 10 1
  timeout 60 $B $mode 0 4
  timeout 60 $B $mode 120 4
  timeout 60 $B $mode 120 6
  timeout 60 $B $mode 60 8 52 11 11
done
timeout 60 $B 0 0 8 104 10 10
timeout 60 $B 2 0 8 104 10 10
} > gpurun_out/fill_probe.log 2>&1
cat gpurun_out/fill_probe.log
