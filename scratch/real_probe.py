"""forward / tap-gradient kernels on real-data-shaped grids: the 64^3 occupancy grid of data-sample/sample_575.npy (tests/golden),
32 copies rolled by random offsets -> clustered grids at 1.6 % overall occupancy"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
g = np.load("tests/golden/vox_sample_575.npz")
occ = np.zeros(64 ** 3, dtype=np.float64); occ[g["restated_count_idx"]] = 1.0
occ = torch.from_numpy(occ.reshape(64, 64, 64))
gen = torch.Generator().manual_seed(0)
xs = []
for s in range(3):
    b = [torch.roll(occ, shifts=(0, int(torch.randint(-15, 15, (1,), generator=gen)), int(torch.randint(-15, 15, (1,), generator=gen))), dims=(0, 1, 2)) for _ in range(32)]
    xs.append(torch.stack(b)[:, None].to(dev))
def t(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
model = bench.kat_model(dev)
K, lam, Kstar, snap = ops.synth_fwd(*bench._spec_params(model))
prep = [ops.prepare(x) for x in xs]
print("state", prep[0][1][:3].tolist(), "occupancy", float((xs[0] != 0).double().mean()))
for od in (torch.float64, torch.float32):
    r = {}
    for name, kw in (("dense", dict(mode=1)), ("mask", dict(mode=2)), ("auto", dict())):
        r[name] = t(lambda i: ops.scenenet_fwd(prep[i % 3][0], Kstar, od, nnz=prep[i % 3][1], **kw))
    print(od, {k: round(v, 1) for k, v in r.items()})
pd = ops.scenenet_fwd(prep[0][0], Kstar, torch.float64, nnz=prep[0][1], mode=1)
pm = ops.scenenet_fwd(prep[0][0], Kstar, torch.float64, nnz=prep[0][1], mode=2)
print("dense vs mask max diff", float((pd - pm).abs().max()), "pred nnz", int((pd > 0).sum()))
g0 = torch.randn(xs[0].shape, device=dev)
for name, kw in (("tapgrad dense", dict(mode=1)), ("tapgrad sparse", dict(mode=2))):
    print(name, round(t(lambda i: ops.tapgrad(prep[i % 3][0], g0, (9, 5, 5), **kw)), 1))
