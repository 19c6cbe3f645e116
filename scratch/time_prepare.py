"""sn_grid_prepare for every input dtype, CUDA-graph replay (memset node + kernel), config-2 batch"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(1)
shape = (32, 1, 64, 64, 64)
base = [(torch.rand(shape, generator=g, device=dev) < 0.016) for _ in range(4)]
def graph_time(fn, n, reps=20):
    s_ = torch.cuda.Stream(device=dev); s_.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s_):
        fn(0)
    torch.cuda.current_stream(dev).wait_stream(s_); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = [fn(i) for i in range(n)]
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps / n * 1e3
for name, xs, nbytes in (("float64", [b.double() for b in base], 8 + 4), ("float32 (count only)", [b.float() for b in base], 4),
                         ("uint8", [b.to(torch.uint8) for b in base], 1 + 4), ("bits", [ops.pack_occupancy(b) for b in base], 0.125 + 4)):
    t = graph_time(lambda i: ops.prepare(xs[i]), 4)
    n = base[0].numel()
    print(f"prepare {name}: {t:.1f} us, {n * nbytes / t / 1e3:.0f} GB/s of algorithmic bytes ({n * nbytes / 1e6:.0f} MB)", flush=True)
    x32, st = ops.prepare(xs[0])
    assert int(st[0]) == int(base[0].sum()) and torch.equal(x32, base[0].float()), name
