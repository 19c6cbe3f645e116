python scratch/prof_fsparse.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:fwd_sparse -s 2 -c 1 -o gpurun_out/prof_fsparse_r1 -f python scratch/prof_fsparse.py > gpurun_out/ncu_fsparse.log 2>&1
tail -2 gpurun_out/ncu_fsparse.log
