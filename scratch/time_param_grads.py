"""the backward's tail: sn_scenenet_param_grads alone (CUDA-graph replay, us)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
model = bench.kat_model(dev)
spec, params = model._spec_and_params()
def graph_time(fn, reps=50, inner=8):
    s_ = torch.cuda.Stream(device=dev); s_.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s_):
        fn()
    torch.cuda.current_stream(dev).wait_stream(s_); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = [fn() for _ in range(inner)]
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps / inner * 1e3
K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params], write_last_lambda=True)
W = torch.randn(spec.kernel_size, device=dev, dtype=torch.float64)
print("synth_fwd", round(graph_time(lambda: ops.synth_fwd(spec, [p.detach() for p in params], write_last_lambda=True)), 2))
print("param_grads", round(graph_time(lambda: ops.param_grads(spec, snap, K, lam, W, 1.0)), 2))
