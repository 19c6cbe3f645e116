python scratch/quantile_probe.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1k.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['training_step']['value'], d['roofline']['frac'], d['roofline']['step_vs_survey_8d_roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'])
"
