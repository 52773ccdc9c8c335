#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list + one full capture of the dominant kernels.
set -x
CMD="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'orb_lap_pc_kernel|orb_bwd_tc_kernel|orb_eval1_l1tc_kernel' -s 30 -c 3 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
