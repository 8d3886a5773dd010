#!/bin/bash
# full ncu capture of the staged tet kernel (second launch of the profile target) -> gpurun_out/prof_tet_staged_$1.ncu-rep
timeout 200 python scripts/profile_target.py C4 3 hot 2>&1 | grep "tet ms" | cut -c1-28 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tet_staged -s 1 -c 1 -f -o gpurun_out/prof_tet_staged_$1 python scripts/profile_target.py C4 3 hot > gpurun_out/ncu_staged_$1.log 2>&1; tail -1 gpurun_out/ncu_staged_$1.log
