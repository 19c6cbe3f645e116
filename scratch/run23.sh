ncu --set full --clock-control none --import-source on -k regex:"stencil_fwd_kernel" -s 2 -c 1 -o gpurun_out/prof_ffma2 -f python scratch/prof_r1.py > gpurun_out/ncu_ffma2.log 2>&1
tail -1 gpurun_out/ncu_ffma2.log
