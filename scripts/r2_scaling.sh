#!/bin/bash
# strong scaling of C4 on N GPUs of one box (and C5 on 8): usage r2_scaling.sh "2 4 8" [c5]
for N in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_n$N.json 2> gpurun_out/r02_bench_c4_n$N.err
  python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([x for x in open("gpurun_out/r02_bench_c4_n%s.json" % n) if x.startswith("{")][-1])
    print("C4 N=%s step %.4f ms value %.4g e2e %.3f ms" % (n, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d["kernel_ms"], "parity", d["parity"] and d["parity"]["ok"])
except Exception as e:
    print("C4 N=%s ERR" % n, e)
PY
done
if [ "$2" = "c5" ]; then
  N=8
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --config C5 --steps 10 --warmup 3 --no-parity > gpurun_out/r02_bench_c5_n$N.json 2> gpurun_out/r02_bench_c5_n$N.err
  tail -c 300 gpurun_out/r02_bench_c5_n$N.json
fi
