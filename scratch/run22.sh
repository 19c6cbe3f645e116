timeout 1200 python -m pytest tests/test_gpu_fwd.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | grep -E "passed|failed|^E " | head
timeout 300 python scratch/time_configs.py 2>&1 | tail -9
