timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -3
python scratch/real_probe.py 2>&1 | grep "torch.float64\|state"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1j.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['training_step']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['fwd_dense']['us'], d['roofline']['step_vs_survey_8d_roofline']['frac'])
print(d['voxelize'])
"
timeout 600 python bench.py --workload config5 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('config5', d['value'], d['ms_per_step'], d['e2e']['value'])"
