timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "128" 2>&1 | tail -25
timeout 300 python scripts/exp/profile_configs.py 2>&1 | tail -5
