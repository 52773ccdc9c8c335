#!/bin/bash
set -x
CMD="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'orb_eval5_tc_kernel' -s 2 -c 1 -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -3
