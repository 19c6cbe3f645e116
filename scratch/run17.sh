timeout 600 python -m pytest tests/test_gpu_peer.py -x -q -s 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_r1c.json 2> gpurun_out/bench_2gpu_r1c.err; tail -3 gpurun_out/bench_2gpu_r1c.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/bench_2gpu_r1c.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['grad_sync_ok'], d['config']['grad_allreduce'])
"
