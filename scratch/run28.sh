timeout 900 python -m pytest tests/test_gpu_criterion.py tests/test_gpu_model.py -m gpu -q 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/b28.err | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['training_step'])" || tail -5 gpurun_out/b28.err
