python scratch/prof_r1.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"stencil_fwd_kernel|tapgrad_sparse_kernel|g0_kernel|prepare_f64|synth_bwd|synth_fwd" -s 12 -c 6 -o gpurun_out/prof_r1_step -f python scratch/prof_r1.py > gpurun_out/ncu_r1_step.log 2>&1
tail -1 gpurun_out/ncu_r1_step.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1b_bench.log 2>&1
tail -1 gpurun_out/ncu_r1b_bench.log | cut -c1-200
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"
