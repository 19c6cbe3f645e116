python scratch/occ_check.py 2>&1 | grep "mismatch\|MISMATCH\|occ=0.016: f64 out dense    9" | head -5
timeout 600 python -m pytest tests/test_gpu_occ.py -m gpu -x -q 2>&1 | tail -2
