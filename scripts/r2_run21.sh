#!/bin/bash
# time line of the end-to-end step for the pipeline orders
for f in 0 3; do
  DSCGPU_E2E_TRACE=1 DSCGPU_PIPELINE_FLAGS=$f timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/r2aa_f$f.json 2> gpurun_out/r2aa_f$f.err
  echo "== flags $f"; grep "dscgpu trace" gpurun_out/r2aa_f$f.err | tail -42 | head -21
done
