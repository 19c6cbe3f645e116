for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${n}gpu_r1i.json 2> gpurun_out/bench_${n}gpu_r1i.err
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_${n}gpu_r1i.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('grad_sync_ok'), d['config'].get('grad_allreduce'), d['e2e']['value'], d['e2e'].get('uint8_occupancy_input',{}).get('value'))
"
done
