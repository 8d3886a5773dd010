#!/bin/bash
for f in 4 2 1; do
DSCGPU_GEOM_LANES=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-overlap-geometry 2>/dev/null | python -c "
import json,sys
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('geom lanes $f:', round(d['ms_per_step'],4), d['kernel_ms'])"
done
