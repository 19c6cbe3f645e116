"""fwd / tap-gradient timing for BASELINE config-4 style shapes (128^3, cubic kernels) and other sizes"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
PEAK = 73.2e12
def t(fn, reps=10):
    for i in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for (B, grid, ks) in [(32, (64,64,64), (9,5,5)), (32, (64,64,64), (9,7,7)), (32, (64,64,64), (6,5,5)), (8, (128,128,128), (9,9,9)), (8, (128,128,128), (11,11,11)),
                      (8, (128,128,128), (13,13,13)), (8, (128,128,128), (15,15,15)), (8, (64,64,256), (9,5,5))]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = (torch.rand((B, 1, *grid), generator=g, device=dev) < 0.016).float()
    K = torch.randn(ks, generator=g, device=dev) * 0.1
    pred = ops.scenenet_fwd(x, K, torch.float32)
    dp = torch.randn(x.shape, generator=g, device=dev)
    g0 = ops.g0(pred, dp)
    T = ks[0] * ks[1] * ks[2]; V = x.numel()
    tf = t(lambda: ops.scenenet_fwd(x, K, torch.float32)); tb = t(lambda: ops.tapgrad(x, g0, ks))
    fl = 2.0 * T * V
    print(f"B={B:2d} grid={grid} k={ks}: fwd {tf*1e6:8.1f} us ({fl/tf/PEAK*100:4.1f}%)  tapgrad {tb*1e6:8.1f} us ({fl/tb/PEAK*100:4.1f}%)  -> {B/(tf+tb):9.0f} grids/s kernels-only")
