#!/bin/bash
# usage (under gpurun --gpus N): scripts/scale_run.sh <tag> <N>
TAG=$1; N=$2
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_scale_1.json 2> gpurun_out/${TAG}_scale_1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_scale_$N.json 2> gpurun_out/${TAG}_scale_$N.err
fi
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${TAG}_scale_$N.json") if l.startswith("{")][-1])
print("N=$N", d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
