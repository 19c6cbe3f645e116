import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dens = float(os.environ.get("DENS", "0.016"))
B, grid, ks = 32, (64, 64, 64), (9, 5, 5)
g = torch.Generator(device=dev).manual_seed(1)
xs = [(torch.rand((B, 1, *grid), generator=g, device=dev) < dens).float() for _ in range(3)]
g0s = [torch.randn(xs[0].shape, generator=g, device=dev) for _ in range(3)]
for i in range(4):
    W = ops.tapgrad(xs[i % 3], g0s[i % 3], ks, mode=2)
torch.cuda.synchronize()
print(float(W.sum()))
