timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_guard.py tests/test_gpu_vmc.py -x -q 2>&1 | tail -15
timeout 300 python scripts/exp/profile_configs.py 2>&1 | tail -8
NQMC_ACT_REUSE=0 timeout 300 python scripts/exp/profile_configs.py 2>&1 | tail -5
