timeout 2000 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -1 gpurun_out/bench_r1i.err | cut -c1-200
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1i.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('uint8_occupancy_input'))
print(d.get('training_step'))
r=d['roofline']; print({k:r[k] for k in ('bound','kernel','achieved','peak','frac','traffic')}); print(r['step_vs_survey_8d_roofline']); print(r['fwd_occupancy_driven'])
print(d['cpu_baseline'], d['clocks'], d['gpu_launches'])
"
ncu --set full --clock-control none --import-source on -k regex:fwd_occ --launch-skip 3 -c 1 -o gpurun_out/prof_fo_r1g -f python scratch/prof_fsparse.py > gpurun_out/ncu_fo_r1g.log 2>&1
