"""does preparing batch i+1 on a second stream while step i runs pay off?  (config 2, CUDA graphs, one B200)
serial: the bench's step (prepare | synthesis -> forward -> G0 -> tap gradient -> row sum -> Jacobian^T), one graph per input set;
piped : prepare graphs on stream P, the rest of the step (model(x, _prepared=...)) on stream M; prepare(i+1) may run beside step i."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
model = bench.kat_model(dev)
params = [p for p in model.parameters() if p.requires_grad]
n_sets = 4
pool = bench.make_pool(dev, 0, n_sets, torch.float64)
modes = ops.select_paths(pool[0][0], (9, 5, 5))
model.path_modes = modes
def capture(fn):
    s_ = torch.cuda.Stream(device=dev); s_.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s_):
        for _ in range(2): out = fn()
    torch.cuda.current_stream(dev).wait_stream(s_); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out
def full_step(x, dp):
    pred = model(x)
    return torch.autograd.grad(pred, params, grad_outputs=dp, allow_unused=True)
serial = [capture(lambda x=x, dp=dp: full_step(x, dp)) for x, dp in pool]
prep = [capture(lambda x=x: ops.prepare(x)) for x, _ in pool]
def main_step(x, dp, pr):
    pred = model(x, _prepared=pr)
    return torch.autograd.grad(pred, params, grad_outputs=dp, allow_unused=True)
mains = [capture(lambda x=x, dp=dp, pr=prep[i][1]: main_step(x, dp, pr)) for i, (x, dp) in enumerate(pool)]
for (gs, outs), (gm, outm), (gp, _) in zip(serial, mains, prep):
    gs.replay(); gp.replay(); gm.replay(); torch.cuda.synchronize()
    for a, b in zip(outs, outm):
        assert (a is None and b is None) or torch.equal(a, b)
K = 200
def time_serial():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(8): serial[i % n_sets][0].replay()
    torch.cuda.synchronize(); a.record()
    for i in range(K): serial[i % n_sets][0].replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / K * 1e3
def time_piped():
    M = torch.cuda.current_stream(dev); P = torch.cuda.Stream(device=dev)
    done = [torch.cuda.Event() for _ in range(K + 10)]
    ready = [torch.cuda.Event() for _ in range(K + 10)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def run(n, timed):
        torch.cuda.synchronize()
        if timed: a.record(M)
        with torch.cuda.stream(P):
            P.wait_stream(M)
            prep[0][0].replay(); ready[0].record(P)
        for i in range(n):
            M.wait_event(ready[i])
            if i + 1 < n:
                with torch.cuda.stream(P):
                    if i >= 1: P.wait_event(done[i - 1])      # at most one step ahead (the buffers of set (i+1) % n_sets are free)
                    prep[(i + 1) % n_sets][0].replay(); ready[i + 1].record(P)
            mains[i % n_sets][0].replay(); done[i].record(M)
        if timed: b.record(M); b.synchronize()
    run(8, False)
    run(K, True)
    return a.elapsed_time(b) / K * 1e3
print("serial us/step", round(time_serial(), 2), round(time_serial(), 2))
print("piped  us/step", round(time_piped(), 2), round(time_piped(), 2))
