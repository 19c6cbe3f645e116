set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/time_configs.py 2>&1 | tail -9
timeout 300 python scratch/time_sparse.py 2>&1 | tail -13
