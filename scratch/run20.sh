P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $P --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_8gpu_r1.json 2> gpurun_out/bench_8gpu_r1.err; echo "rc=$?"
SN_BENCH_NCCL=1 timeout 600 $P --master-port 29532 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_8gpu_nccl_r1.json 2> gpurun_out/bench_8gpu_nccl_r1.err; echo "rc=$?"
python - <<'PY'
import json
for f in ["bench_8gpu_r1", "bench_8gpu_nccl_r1"]:
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["uint8_occupancy_input"]["value"], d["grad_sync_ok"], d["config"]["grad_allreduce"])
PY
grep -v "Warning\|run_backward\|^$" gpurun_out/bench_8gpu_r1.err | tail -5 | cut -c1-250
