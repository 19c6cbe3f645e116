timeout 600 python scratch/occ_check.py > gpurun_out/occ_check4.log 2>&1; tail -11 gpurun_out/occ_check4.log
timeout 600 python -m pytest tests/test_gpu_occ.py tests/test_gpu_fwd.py -m gpu -x -q 2>&1 | tail -3
