"""kernel-only timing of fwd / bwd at config-2 size with CUDA events (rotating inputs)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
io = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
model = bench.kat_model(dev)
pool = bench.make_pool(dev, 0, 4, io)
x32s = [ops.cast_f32(p[0]) for p in pool]
K, lam, Kstar, snap = ops.synth_fwd(*bench._spec_params(model))
preds = [ops.scenenet_fwd(x, Kstar, io) for x in x32s]
def t(fn, reps=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
tf = t(lambda i: ops.scenenet_fwd(x32s[i % 4], Kstar, io))
tb = t(lambda i: ops.scenenet_bwd(x32s[i % 4], preds[i % 4], pool[i % 4][1], bench.KERNEL))
print(f"io={sys.argv[1] if len(sys.argv)>1 else 'f64'} stagger={os.environ.get('SN_BWD_STAGGER_NS','0')}: fwd {tf:.1f} us ({51.57/tf*100:.1f}% of 73.2TF roofline)  bwd(g0+main+reduce) {tb:.1f} us")
