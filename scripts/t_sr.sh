python -m pytest tests/test_gpu_sr.py -x -q 2>&1 | tail -25
