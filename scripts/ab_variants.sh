#!/bin/bash
# A/B of tet-kernel variants on the bench workload: bash scripts/ab_variants.sh 128 8320 16512 ...
for v in "$@"; do
    DSCGPU_TET_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_v$v.json 2> gpurun_out/b_v$v.err
    python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads([x for x in open("gpurun_out/b_v%s.json" % v) if x.startswith("{")][-1])
    print("variant", v, "step", round(d["ms_per_step"], 4), d["kernel_ms"], "e2e", round(d["e2e"]["ms_per_step"], 3), d["check"]["mean"])
except Exception as e:
    print("variant", v, "ERR", e)
PY
done
