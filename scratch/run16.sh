python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu_r1b.json 2> gpurun_out/bench_2gpu_r1b.err; tail -2 gpurun_out/bench_2gpu_r1b.err | cut -c1-300; python -c "
import json
for l in open('gpurun_out/bench_2gpu_r1b.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['grad_sync_ok'])
"
