#!/bin/bash
# Round-2 evidence on one B200: launch list of the default bench command, full captures of the staged tet kernel, the exact-list
# kernel and the fused force exchange kernel.  Outputs under gpurun_out/.
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tet_staged -s 1 -c 1 -f -o gpurun_out/prof_tet_staged_r02 python scripts/profile_target.py C4 3 hot > gpurun_out/r02_ncu_staged.log 2>&1; tail -1 gpurun_out/r02_ncu_staged.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tet_exact_list -s 1 -c 1 -f -o gpurun_out/prof_tet_exact_r02 python scripts/profile_target.py C4 3 hot > gpurun_out/r02_ncu_exact.log 2>&1; tail -1 gpurun_out/r02_ncu_exact.log
PART=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:comm_force_fused -s 1 -c 1 -f -o gpurun_out/prof_comm_force_r02 python scripts/profile_comm.py C4 3 > gpurun_out/r02_ncu_comm.log 2>&1; tail -1 gpurun_out/r02_ncu_comm.log
PART=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_comm_launches_part8.csv python scripts/profile_comm.py C4 3 > /dev/null 2>&1
grep "comm_\|force_geom\|tet_staged\|tet_exact" gpurun_out/r02_comm_launches_part8.csv | awk -F'","' '{print $5, $NF}' | tail -8
