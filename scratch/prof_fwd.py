"""profiling target: the two forward kernels on the config-2 batch (KAT kernel); also prints CUDA-graph-replay timings (no host gaps)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import ops
from oracle import model_oracle as mo, ref_shim
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dens = float(os.environ.get("DENS", "0.016"))
torch.manual_seed(0)
m = sb.SceneNet(dict(mo.KAT_GENEO_NUM), (9, 5, 5)).to(dev)
ref_shim.set_scenenet_params(m, mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST)
spec, params = m._spec_and_params()
K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params])
g = torch.Generator(device=dev).manual_seed(1)
xs = [(torch.rand((32, 1, 64, 64, 64), generator=g, device=dev) < dens).double() for _ in range(3)]
prep = [ops.prepare(x) for x in xs]
for i in range(4):
    pd = ops.scenenet_fwd(prep[i % 3][0], Kstar, torch.float64, nnz=prep[i % 3][1], mode=1)
    ps = ops.scenenet_fwd(prep[i % 3][0], Kstar, torch.float64, nnz=prep[i % 3][1], mode=2)
torch.cuda.synchronize()
print(float(pd.sum()), float(ps.sum()), float((pd - ps).abs().max()))
if os.environ.get("GRAPH_TIMES"):
    def graph_time(fn, reps=20):
        s_ = torch.cuda.Stream(device=dev); s_.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s_):
            fn(0)
        torch.cuda.current_stream(dev).wait_stream(s_); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            outs = [fn(i) for i in range(3)]
        gr.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): gr.replay()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / reps / 3 * 1e3
    for d in (0.0, 0.005, 0.016, 0.03, 0.05, 0.08):
        xs2 = [(torch.rand((32, 1, 64, 64, 64), generator=g, device=dev) < d).double() for _ in range(3)]
        pr2 = [ops.prepare(x) for x in xs2]
        res = {}
        for mode, tag in ((1, "dense"), (2, "sparse"), (0, "auto")):
            for odt, on in ((torch.float64, "f64"), (torch.float32, "f32")):
                res[f"{tag}/{on}"] = round(graph_time(lambda i: ops.scenenet_fwd(pr2[i][0], Kstar, odt, nnz=pr2[i][1], mode=mode)), 1)
        print("graph-replay us, occupancy", d, res, flush=True)
    z = torch.empty(32 * 64 ** 3, dtype=torch.float64, device=dev)
    print("memset 67 MB us:", round(graph_time(lambda i: z.zero_()), 1))
