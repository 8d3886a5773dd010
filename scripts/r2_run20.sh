#!/bin/bash
# e2e step: pipeline flags A/B with the event-timer breakdown
for f in 0 1 2 3; do for c in 4; do
  DSCGPU_PIPELINE_FLAGS=$f DSCGPU_PIPELINE_CHUNKS=$c timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/r2z_bench_f${f}_c$c.json 2> gpurun_out/r2z_bench_f${f}_c$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2z_bench_f${f}_c$c.json'))
print('flags $f chunks $c', 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['parts'])
PY
done; done
DSCGPU_PIPELINE_FLAGS=0 DSCGPU_PIPELINE_CHUNKS=2 timeout 600 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags 0 chunks 2', round(d['e2e']['ms_per_step'],4), d['e2e']['parts'])"
DSCGPU_PIPELINE_FLAGS=0 DSCGPU_PIPELINE_CHUNKS=0 timeout 600 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags 0 chunks 0', round(d['e2e']['ms_per_step'],4), d['e2e']['parts'])"
