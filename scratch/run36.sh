timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -3 gpurun_out/bench_r1g.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1g.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('uint8_occupancy_input'))
print(d.get('training_step'))
r=d['roofline']; print({k:r[k] for k in ('bound','kernel','achieved','peak','frac','traffic')}); print(r['fwd_dense'], r['fwd_occupancy_driven']); print(r['prepare_pass'], r['g0_pass'], r['bwd_tapgrad_occupancy_driven']['us'])
"
