timeout 600 python -m pytest tests/test_gpu_tapgrad.py tests/test_gpu_model.py -m gpu -q 2>&1 | grep -E "passed|failed|^E " | head -5
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['roofline']['prepare_pass'])"
