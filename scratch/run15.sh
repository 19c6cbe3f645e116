timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|^E " | head
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_occupancy_input'], d['config']['eager_module_value'])
print(d['training_step'])
"
