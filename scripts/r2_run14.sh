#!/bin/bash
for o in "" "--no-overlap-geometry"; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity $o 2>/dev/null | python -c "
import json,sys
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('opt=[$o]', round(d['ms_per_step'],4), d['kernel_ms'], 'e2e', round(d['e2e']['ms_per_step'],3), 'incr', round(d['e2e_incremental']['ms_per_step'],3), d['with_sumsq']['ms_per_step'])"
done
