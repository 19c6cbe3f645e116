timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "synthetic_fixed or random_params" 2>&1 | grep -E "worst|passed|failed|Assertion" | head -30
echo "---- tanh64"
SN_FS_TANH64=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "synthetic_fixed or random_params" 2>&1 | grep -E "worst|passed|failed|Assertion" | head -30
echo "---- dense fwd"
SN_SPARSE_FWD_PCT=0 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -s -k "synthetic_fixed or random_params" 2>&1 | grep -E "worst|passed|failed|Assertion" | head -30
timeout 300 python scratch/vox_bench.py 2>&1 | tail -6
