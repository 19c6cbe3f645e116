import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
import scenenet_b200 as sb
from scenenet_b200 import ops, voxel_ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
SCANS, NPTS, GRID_XYZ = 8, 120_000, (64, 64, 256)
g = torch.Generator().manual_seed(77)
r = 5.0 + 45.0 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64) ** 2
th = 2 * 3.141592653589793 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64)
z = -3.0 + 6.0 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64) ** 3
lab = torch.where(torch.rand(SCANS * NPTS, generator=g) < 0.01, 80.0, 40.0).to(torch.float64)
rows = torch.stack([(r * torch.cos(th)).float().double(), (r * torch.sin(th)).float().double(), z.float().double(), lab], 1).contiguous().to(dev)
off = (torch.arange(0, SCANS + 1, dtype=torch.int64) * NPTS).to(dev)
def t(fn, reps=20):
    for i in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
vox = lambda: voxel_ops.voxelize_clouds(rows[:, :3], off, GRID_XYZ, rows[:, 3], [80.0], want=("occ",), occ_dtype=torch.float32)
x = vox()["occ"].view(SCANS, 1, 256, 64, 64)
print("occupancy", float((x != 0).float().mean()), "per z-tile layer:", [round(float((x[:, :, 8*i:8*i+8] != 0).float().mean()), 3) for i in range(0, 32, 4)])
model = bench.kat_model(dev)
K, lam, Kstar, snap = ops.synth_fwd(*bench._spec_params(model))
x32, st = ops.prepare(x)
print("voxelize", t(vox), "prepare", t(lambda: ops.prepare(x)))
for name, kw in (("dense", dict(mode=1)), ("mask", dict(nnz=st, mode=2)), ("scan", dict(mode=2)), ("auto", dict(nnz=st))):
    print(name, t(lambda: ops.scenenet_fwd(x32, Kstar, torch.float32, **kw)))
with torch.no_grad():
    print("model fwd eager", t(lambda: model(x)))
g0 = torch.randn(x.shape, device=dev)
for name, kw in (("tapgrad dense", dict(mode=1)), ("tapgrad sparse", dict(mode=2)), ("tapgrad auto", dict(nnz=st))):
    print(name, t(lambda: ops.tapgrad(x32, g0, (9, 5, 5), **kw)))
print("state", st[:3].tolist())
