set -x
python scratch/prof_fsparse.py > gpurun_out/plain_fs.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fwd_sparse --launch-skip 3 -c 1 -o gpurun_out/prof_fs_r1c -f python scratch/prof_fsparse.py > gpurun_out/ncu_fs_r1c.log 2>&1
DENS=0.0 ncu --set full --clock-control none --import-source on -k regex:fwd_sparse --launch-skip 3 -c 1 -o gpurun_out/prof_fs0_r1c -f python scratch/prof_fsparse.py > gpurun_out/ncu_fs0_r1c.log 2>&1
timeout 300 python scratch/time_sparse.py 2>&1 | tail -30
