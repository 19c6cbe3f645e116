python scratch/real_probe.py 2>&1 | tail -8
python scratch/occ_check.py 2>&1 | grep "occ=0.000\|occ=0.016: f64 out dense" | head -3
timeout 900 python -m pytest tests/test_gpu_fwd.py tests/test_gpu_occ.py tests/test_gpu_model.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
