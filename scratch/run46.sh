for k in 9 15; do timeout 600 python bench.py --workload config4 --kernel $k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c4_${k}_r1i.json 2> gpurun_out/c4_${k}_r1i.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c4_${k}_r1i.json') if l.startswith('{')][-1])
r=d['roofline']; print('config4 k=$k', d['value'], d['ms_per_step'], r['kernel'], r['fwd_dense']['us'], r['fwd_occupancy_driven']['us'], r['bwd_tapgrad_occupancy_driven']['us'])
"; done
timeout 600 python bench.py --workload config5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c5_r1i.json 2> gpurun_out/c5_r1i.err; tail -2 gpurun_out/c5_r1i.err | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/c5_r1i.json') if l.startswith('{')][-1])
print('config5', d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'])
"
