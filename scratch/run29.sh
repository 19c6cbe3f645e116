timeout 900 python -m pytest tests/test_gpu_voxel.py -m gpu -q 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline())
for k,v in d['voxelize'].items():
    if k!='note': print(k, {a:round(b,1) for a,b in v.items()})"
