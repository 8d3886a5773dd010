#!/bin/bash
for fcr in "3 3 1" "3 4 1" "3 5 1" "7 4 0.85" "7 4 0.75" "7 3 0.7" "3 4 1"; do set -- $fcr
  DSCGPU_PIPELINE_FLAGS=$1 DSCGPU_PIPELINE_CHUNKS=$2 DSCGPU_PIPELINE_RATIO=$3 timeout 600 python bench.py --no-cpu-baseline --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags $1 chunks $2 ratio $3', round(d['e2e']['ms_per_step'],4), {k:v for k,v in d['e2e']['parts'].items() if k!='what'})"
done
