timeout 900 python -m pytest tests/test_gpu_occ.py tests/test_gpu_fwd.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -12
python scratch/occ_check.py 2>&1 | grep "mismatches\|occ=0.016: f64 out dense    9"
