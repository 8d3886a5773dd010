#!/bin/bash
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_geom -s 1 -c 1 -f -o gpurun_out/prof_force_geom_r02 python scripts/profile_target.py C4 3 force > gpurun_out/r02_ncu_fgeom.log 2>&1; tail -1 gpurun_out/r02_ncu_fgeom.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_node2 -s 1 -c 1 -f -o gpurun_out/prof_force_node2_r02 python scripts/profile_target.py C4 3 force > gpurun_out/r02_ncu_fnode.log 2>&1; tail -1 gpurun_out/r02_ncu_fnode.log
