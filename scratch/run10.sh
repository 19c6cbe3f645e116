timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -3
timeout 300 python scratch/vox_bench.py 2>&1 | tail -14
