python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b51.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu51.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo_r1h -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo_r1h.log 2>&1
ncu --set full --clock-control none -k regex:prepare_f64 --launch-skip 1 -c 1 -o gpurun_out/prof_prep_r1h -f python scratch/prof_fsparse.py > gpurun_out/ncu_prep_r1h.log 2>&1
ls -la gpurun_out/r1c_launches_bench.csv
