set -x
timeout 1200 python -m pytest tests/test_gpu_fwd.py -x -q 2>&1 | tail -15
timeout 300 python scratch/time_sparse.py 2>&1 | tail -10
