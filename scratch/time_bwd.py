"""backward on point-cloud grids: fused G0 + tap gradient (sn_scenenet_bwd with a state buffer) against g0_kernel +
tapgrad_sparse_kernel, CUDA-graph replay over rotating inputs (> L2), config-2 batch"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
from scenenet_b200._lib import SN_TAPGRAD_AUTO, SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
g = torch.Generator(device=dev).manual_seed(1)
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
def run(shape, ks, dens, dt, n=4):
    xs = [(torch.rand(shape, generator=g, device=dev) < dens).float() for _ in range(n)]
    prep = [ops.prepare(x) for x in xs]
    pred = [torch.relu(torch.tanh(torch.randn(shape, generator=g, device=dev, dtype=dt))) for _ in range(n)]
    dpred = [torch.randn(shape, generator=g, device=dev, dtype=dt) for _ in range(n)]
    out = {}
    out["fused"] = graph_time(lambda i: ops.scenenet_bwd(prep[i][0], pred[i], dpred[i], ks, nnz=prep[i][1], mode=SN_TAPGRAD_SPARSE), n)
    out["auto"] = graph_time(lambda i: ops.scenenet_bwd(prep[i][0], pred[i], dpred[i], ks, nnz=prep[i][1], mode=SN_TAPGRAD_AUTO), n)
    out["g0+sparse"] = graph_time(lambda i: ops.scenenet_bwd(prep[i][0], pred[i], dpred[i], ks, mode=SN_TAPGRAD_SPARSE), n)
    g0s = [ops.g0(pred[i], dpred[i]) for i in range(n)]
    out["ring"] = graph_time(lambda i: ops.tapgrad(prep[i][0], g0s[i], ks, nnz=prep[i][1], mode=SN_TAPGRAD_SPARSE), n)
    out["tiles"] = graph_time(lambda i: ops.tapgrad(prep[i][0], g0s[i], ks, mode=SN_TAPGRAD_SPARSE), n)
    out["g0"] = graph_time(lambda i: ops.g0(pred[i], dpred[i]), n)
    print(shape, ks, dens, str(dt).split(".")[-1], {k: round(v, 1) for k, v in out.items()}, flush=True)
for dens in (0.016, 0.0, 0.005, 0.03, 0.05, 0.08):
    run((32, 1, 64, 64, 64), (9, 5, 5), dens, torch.float64)
run((32, 1, 64, 64, 64), (9, 5, 5), 0.016, torch.float32)
run((4, 1, 128, 128, 128), (9, 9, 9), 0.016, torch.float64, n=2)
run((4, 1, 128, 128, 128), (15, 15, 15), 0.016, torch.float64, n=2)
