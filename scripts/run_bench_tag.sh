#!/bin/bash
# usage (under gpurun): scripts/run_bench_tag.sh <tag>  -> gpurun_out/<tag>_bench.json
python bench.py > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err
tail -c 300 gpurun_out/$1_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/$1_bench.json"))
print(d["value"], d["e2e"]["value"], d["ms_per_step"])
print(d.get("sr_iteration"))
PY
