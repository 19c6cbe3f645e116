timeout 600 python scratch/occ_check.py > gpurun_out/occ_check.log 2>&1; tail -12 gpurun_out/occ_check.log
ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo_r1e -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo_r1e.log 2>&1
