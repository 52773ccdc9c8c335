#!/bin/bash
# usage: ncu_k.sh <kernel regex> <out name> ; full capture of one launch of the matching kernel
CMD="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s 2 -c 1 -o gpurun_out/$2 $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/$2.ncu-rep
