python scratch/c5_probe.py 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_occ.py tests/test_gpu_fwd.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload config5 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('config5', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('config2', d['value'], d['ms_per_step'], d['roofline']['kernel'])"
