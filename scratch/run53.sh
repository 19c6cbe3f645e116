timeout 900 python -m pytest tests/test_gpu_criterion.py tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['training_step'])"
