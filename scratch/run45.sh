SEED=11 N=300 timeout 900 python scratch/fuzz_kernels.py 2>&1 | tail -5
SEED=12 N=300 timeout 900 python scratch/fuzz_kernels.py 2>&1 | tail -5
