python bench.py --workload config4 --kernel 15 --steps 10 --warmup 3 > gpurun_out/c4_15.json 2> gpurun_out/c4_15.err; echo "rc=$?"
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$P --master-port 29521 bench.py --gpus 2 --workload config4 --kernel 9 --steps 10 --warmup 3 > gpurun_out/c4_9_2gpu.json 2> gpurun_out/c4_9_2gpu.err; echo "rc=$?"
$P --master-port 29522 bench.py --gpus 2 --workload config5 --steps 20 --warmup 3 > gpurun_out/c5_2gpu.json 2> gpurun_out/c5_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
for f in ["c4_15", "c4_9_2gpu", "c5_2gpu"]:
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l); r = d.get("roofline") or {}
            print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("fwd"), (r.get("bwd_tapgrad_dense") or {}).get("us"), (r.get("bwd_tapgrad_occupancy_driven") or {}).get("us"), r.get("fwd_occupancy_driven"))
PY
