timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^FAILED|^E  " | head -12
python scratch/prof_vox.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:bin_kernel -s 2 -c 1 -o gpurun_out/prof_bin_r1 -f python scratch/prof_vox.py > gpurun_out/ncu_bin.log 2>&1
tail -2 gpurun_out/ncu_bin.log
