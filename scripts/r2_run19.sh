#!/bin/bash
# pipelined upload reordered (faces after chunk 0, CSR hoisted, exact list per chunk): tests, then e2e for 4 / 6 / 8 chunks
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1; tail -3 gpurun_out/r2y_tests.log
for c in 4 6 8; do
  DSCGPU_PIPELINE_CHUNKS=$c timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2y_bench_c$c.json 2> gpurun_out/r2y_bench_c$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench_c$c.json'))
print('chunks $c', 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'incr', round(d['e2e_incremental']['ms_per_step'],4), 'parity', d.get('parity',{}).get('ok'))
PY
done
