python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_r1h.json 2> gpurun_out/bench_2gpu_r1h.err; tail -2 gpurun_out/bench_2gpu_r1h.err | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu_r1h.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['n_gpus'], d.get('grad_sync_ok'), d['config'].get('grad_allreduce'), d['e2e']['value'])
"
timeout 300 python -m pytest tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -2
