python -m pytest tests/test_gpu_sr.py tests/test_gpu_vmc.py -x -q 2>&1 | tail -25
