timeout 600 python scratch/occ_check.py > gpurun_out/occ_check2.log 2>&1; tail -12 gpurun_out/occ_check2.log
timeout 900 python -m pytest tests/test_gpu_occ.py tests/test_gpu_metrics.py tests/test_gpu_fwd.py -m gpu -x -q 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo_r1f -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo_r1f.log 2>&1
