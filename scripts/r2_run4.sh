#!/bin/bash
# full GPU tests, bench A/B of w2 against the staged kernel (C4, then C3 and C2), launch list
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; tail -3 gpurun_out/r2d_tests.log
bash scripts/ab_variants.sh 16512 147584
for c in C3 C2; do for v in 16512 147584; do
  DSCGPU_TET_VARIANT=$v timeout 200 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$c', $v, round(d['ms_per_step'],4), d['kernel_ms'], d['roofline']['frac'])"
done; done
