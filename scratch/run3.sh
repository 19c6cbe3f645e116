set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scratch/time_sparse.py 2>&1 | tail -14
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -3 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
