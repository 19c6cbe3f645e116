import os, sys, torch
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
torch.manual_seed(0)
qnet = sb.SCENENetQuantile({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), qs=torch.tensor([0.1, 0.5, 0.9]), device=dev)
g = torch.Generator(device=dev).manual_seed(1)
xs = [(torch.rand((32, 1, 64, 64, 64), generator=g, device=dev) < 0.016).double() for _ in range(3)]
def t(fn, reps=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
per = t(lambda i: qnet(xs[i % 3]))
with torch.no_grad():
    fused = t(lambda i: qnet(xs[i % 3]))
prep = [ops.prepare(x) for x in xs]
Ks = torch.randn((3, 9, 5, 5), generator=g, device=dev) * 0.1
k3 = t(lambda i: ops.scenenet_fwd_multi(prep[i % 3][0], Ks, torch.float64, nnz=prep[i % 3][1], mode=2))
k1 = t(lambda i: [ops.scenenet_fwd(prep[i % 3][0], Ks[q], torch.float64, nnz=prep[i % 3][1], mode=2) for q in range(3)])
print(f"SCENENetQuantile (3 observers, B=32, 64^3): per-observer module path {per:.1f} us, fused inference {fused:.1f} us; kernels: 3 launches {k1:.1f} us, one multi launch {k3:.1f} us")
