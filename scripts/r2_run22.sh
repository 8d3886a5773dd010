#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_tests.log 2>&1; tail -3 gpurun_out/r2ab_tests.log
for fc in "0 4" "3 4" "7 4" "7 5" "7 6"; do set -- $fc
  DSCGPU_PIPELINE_FLAGS=$1 DSCGPU_PIPELINE_CHUNKS=$2 timeout 600 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags $1 chunks $2', round(d['e2e']['ms_per_step'],4), 'incr', round(d['e2e_incremental']['ms_per_step'],4), {k:v for k,v in d['e2e']['parts'].items() if k!='what'})"
done
DSCGPU_E2E_TRACE=1 DSCGPU_PIPELINE_FLAGS=7 DSCGPU_PIPELINE_CHUNKS=5 timeout 600 python bench.py --no-cpu-baseline --no-parity > /dev/null 2> gpurun_out/r2ab_trace.err
L=$(grep -n "start (copy" gpurun_out/r2ab_trace.err | tail -2 | head -1 | cut -d: -f1); sed -n "$((L-1)),$((L+20))p" gpurun_out/r2ab_trace.err
