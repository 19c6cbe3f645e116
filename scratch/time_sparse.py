"""dense vs occupancy-driven tap gradient: timing over occupancy and kernel sizes"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
def t(fn, reps=10):
    for i in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
cases = [(32, (64,64,64), (9,5,5), d) for d in (0.0, 0.016, 0.05, 0.1, 0.2, 0.3, 0.5, 1.0)]
cases += [(32, (64,64,64), (9,7,7), 0.016), (8, (128,128,128), (9,9,9), 0.016), (8, (128,128,128), (15,15,15), 0.016), (8, (64,64,256), (9,5,5), 0.016)]
for (B, grid, ks, dens) in cases:
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [(torch.rand((B, 1, *grid), generator=g, device=dev) < dens).float() for _ in range(3)]
    g0s = [torch.randn(xs[0].shape, generator=g, device=dev) for _ in range(3)]
    i = [0]
    def run(mode):
        i[0] += 1
        return ops.tapgrad(xs[i[0] % 3], g0s[i[0] % 3], ks, mode=mode)
    td = t(lambda: run(1)); ts = t(lambda: run(2))
    x32, nnz = ops.prepare(xs[0])
    ta = t(lambda: ops.tapgrad(xs[0], g0s[0], ks, nnz=nnz, mode=0))
    tp = t(lambda: ops.prepare(xs[i[0] % 3]))
    print(f"B={B:2d} grid={grid} k={ks} occ={dens:5.3f}: dense {td*1e6:8.1f} us  sparse {ts*1e6:8.1f} us  auto {ta*1e6:8.1f} us  prepare(f32 count) {tp*1e6:6.1f} us", flush=True)
print("---- forward", flush=True)
fcases = [(32, (64,64,64), (9,5,5), d) for d in (0.0, 0.016, 0.03, 0.05, 0.1, 0.3)]
fcases += [(32, (64,64,64), (9,7,7), 0.016), (8, (128,128,128), (9,9,9), 0.016), (8, (64,64,256), (9,5,5), 0.016)]
for (B, grid, ks, dens) in fcases:
    g = torch.Generator(device=dev).manual_seed(1)
    xs = [(torch.rand((B, 1, *grid), generator=g, device=dev) < dens).float() for _ in range(3)]
    K = torch.randn(ks, generator=g, device=dev) * 0.1
    i = [0]
    def runf(mode, dt=torch.float64):
        i[0] += 1
        return ops.scenenet_fwd(xs[i[0] % 3], K, dt, mode=mode)
    td = t(lambda: runf(1)); ts = t(lambda: runf(2))
    print(f"B={B:2d} grid={grid} k={ks} occ={dens:5.3f}: fwd dense {td*1e6:8.1f} us  sparse {ts*1e6:8.1f} us", flush=True)
