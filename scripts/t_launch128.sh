python scripts/exp/launches_h128.py > gpurun_out/l128_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_h128_launches.csv python scripts/exp/launches_h128.py > gpurun_out/l128_ncu.log 2>&1
tail -3 gpurun_out/l128_ncu.log
