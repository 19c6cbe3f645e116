timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q 2>&1 | grep -E "passed|failed|^E " | head -8
