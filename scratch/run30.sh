timeout 900 python -m pytest tests/test_gpu_tapgrad.py tests/test_gpu_model.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tapgrad_sparse" -c 8 --csv python scratch/prof_r1.py 2>/dev/null | grep tapgrad | awk -F, '{print $NF}' | tr '\n' ' '; echo
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'])"
