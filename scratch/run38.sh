timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_metrics.py -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; tail -2 gpurun_out/bench_r1h.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1h.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step']); print(d['voxelize'].get('device_loader_32x60k_pts_64^3'))
"
