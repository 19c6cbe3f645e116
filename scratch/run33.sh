timeout 600 python scratch/occ_check.py 2>&1 | tail -14
ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo_r1d -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo_r1d.log 2>&1
DENS=0.0 ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo0_r1d -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo0_r1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prepare_f64 --launch-skip 1 -c 1 -o gpurun_out/prof_prep_r1d -f python scratch/prof_fsparse.py > gpurun_out/ncu_prep_r1d.log 2>&1
