set -x
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -5 gpurun_out/bench_r1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1c.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['uint8_occupancy_input'])
print(d['training_step'])
print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!='note'}) for k,v in d['roofline'].items() if k in ('fwd','bwd_tapgrad_dense','bwd_tapgrad_occupancy_driven','g0_pass','prepare_pass','frac')})
"
