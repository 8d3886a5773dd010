#!/bin/bash
for k in 4 3 2; do
DSCGPU_EXACT_CTAS_PER_SM=$k timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('exact CTAs/SM $k:', round(d['ms_per_step'],4), d['kernel_ms'])"
done
