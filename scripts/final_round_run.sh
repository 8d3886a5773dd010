#!/bin/bash
# End-of-round evidence run on one B200: GPU tests, smoke, the default bench line, the ncu launch list of the same command and one
# full capture of the dominant kernel.  Outputs under gpurun_out/.
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final2_tests.log 2>&1; tail -1 gpurun_out/final2_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log
timeout 300 python bench.py > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; tail -c 400 gpurun_out/final2_bench_n1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final2_ncu_list.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tet_integrate_w2 -s 1 -c 1 -f -o gpurun_out/prof_tet_w2_r01final python scripts/profile_target.py C4 3 hot > gpurun_out/final2_ncu_full.log 2>&1; tail -2 gpurun_out/final2_ncu_full.log
