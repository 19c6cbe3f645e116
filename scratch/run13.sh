timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "worst|passed|failed|Assertion|^E " | head -40
timeout 300 python scratch/time_sparse.py 2>&1 | tail -10
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1e.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_occupancy_input'])
print(d['training_step'])
for k,v in d['voxelize'].items(): print(k, v)
r=d['roofline']; print(r['fwd'], r['bwd_tapgrad_occupancy_driven']['us'], r['prepare_pass'])
"
