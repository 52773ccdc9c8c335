#!/bin/bash
# usage (under gpurun): scripts/ncu_r2.sh <tag> [kernel regex]
# bench line, ncu launch list and one --set full capture of the dominant kernels, all into gpurun_out/<tag>_*
TAG=${1:-r2}
KRE=${2:-'orb_lap_pc_kernel|orb_bwd_tc_kernel|orb_eval1_l1tc_kernel'}
CMD="python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu-baseline --no-extras"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
for K in $(echo "$KRE" | tr '|' ' '); do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o gpurun_out/${TAG}_prof_$K $CMD > gpurun_out/${TAG}_ncu_full_$K.log 2>&1
done
ls -la gpurun_out | tail -8
# the TMA-fed kernels: SR SYRK (scripts/time_sr.py) and the 128-wide reverse pass (scripts/exp/launches_h128.py)
python scripts/time_sr.py > gpurun_out/${TAG}_sr_timing.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 -o gpurun_out/${TAG}_prof_gram python scripts/time_sr.py > gpurun_out/${TAG}_ncu_full_gram.log 2>&1
python scripts/exp/launches_h128.py > /dev/null 2>&1 &&
for K in bwd128_d1_kernel bwd128_gram_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o gpurun_out/${TAG}_prof_$K python scripts/exp/launches_h128.py > gpurun_out/${TAG}_ncu_full_$K.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
