timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py tests/test_gpu_vmc.py -x -q 2>&1 | tail -5
python bench.py --no-extras --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python -c "
import json;d=json.load(open('gpurun_out/q_bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline_context']['kernel_ms'])"
