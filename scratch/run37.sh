timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -8
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b37.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu37.log 2>&1
ls -la gpurun_out/r1c_launches_bench.csv
