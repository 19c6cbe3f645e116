import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
from scenenet_b200._lib import SN_TAPGRAD_SPARSE
dev = torch.device("cuda", 0)
dens = float(sys.argv[1]) if len(sys.argv) > 1 else 0.016
g = torch.Generator(device=dev).manual_seed(1)
shape = (32, 1, 64, 64, 64)
x = (torch.rand(shape, generator=g, device=dev) < dens).float()
x32, st = ops.prepare(x)
g0 = torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
for _ in range(3):
    W = ops.tapgrad(x32, g0, (9, 5, 5), nnz=st, mode=SN_TAPGRAD_SPARSE)
torch.cuda.synchronize()
W2 = ops.tapgrad(x32, g0, (9, 5, 5), mode=SN_TAPGRAD_SPARSE)
print("max diff", float((W - W2).abs().max()), float(W2.abs().max()))
