import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dens = float(os.environ.get("DENS", "0.016"))
B, grid, ks = 32, (64, 64, 64), (9, 5, 5)
g = torch.Generator(device=dev).manual_seed(1)
xs = [(torch.rand((B, 1, *grid), generator=g, device=dev) < dens).float() for _ in range(3)]
K = torch.randn(ks, generator=g, device=dev) * 0.1
prep = [ops.prepare(x.double()) for x in xs]
odt = torch.float32 if os.environ.get("OUT32") else torch.float64
for i in range(4):
    p = ops.scenenet_fwd(prep[i % 3][0], K, odt, nnz=prep[i % 3][1], mode=2)
torch.cuda.synchronize()
print(float(p.sum()))
