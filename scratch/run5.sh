set -x
timeout 1200 python -m pytest tests/test_gpu_criterion.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15
