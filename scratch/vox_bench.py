"""voxelization throughput sweep (BASELINE config 3)"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import voxel_ops
dev = torch.device("cuda", 0)

def cloud(n, seed=0, coherent=True):
    g = torch.Generator(device=dev).manual_seed(seed)
    n_g, n_v = int(0.70 * n), int(0.25 * n); n_t = n - n_g - n_v
    ground = torch.stack([torch.rand(n_g, generator=g, device=dev, dtype=torch.float64) * 30, torch.rand(n_g, generator=g, device=dev, dtype=torch.float64) * 30,
                          torch.randn(n_g, generator=g, device=dev, dtype=torch.float64) * 0.3], 1)
    cent = torch.rand((20, 3), generator=g, device=dev, dtype=torch.float64) * torch.tensor([30, 30, 9.0], device=dev, dtype=torch.float64)
    veg = cent[torch.randint(0, 20, (n_v,), generator=g, device=dev)] + torch.randn((n_v, 3), generator=g, device=dev, dtype=torch.float64) * 2.0
    tower = torch.stack([torch.randn(n_t, generator=g, device=dev, dtype=torch.float64) * 0.5 + 15, torch.randn(n_t, generator=g, device=dev, dtype=torch.float64) * 0.5 + 15,
                         torch.rand(n_t, generator=g, device=dev, dtype=torch.float64) * 40], 1)
    pts = torch.cat([ground, veg, tower]) + torch.tensor([544850.0, 4634550.0, 160.0], device=dev, dtype=torch.float64)
    lab = torch.cat([torch.randint(1, 13, (n_g + n_v,), generator=g, device=dev).double(), torch.full((n_t,), 15.0, device=dev, dtype=torch.float64)])
    if coherent:   # scan-coherent order: sorted by a coarse 2 m tile id
        key = ((pts[:, 0] - 544850.0) / 2).floor() * 64 + ((pts[:, 1] - 4634550.0) / 2).floor()
        order = torch.argsort(key)
    else:
        order = torch.randperm(n, generator=g, device=dev)
    rows = torch.cat([pts[order], lab[order, None]], 1).contiguous()   # [N,4] like the TS40K npy
    return rows

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e-3

for n in [100_000, 1_000_000, 10_000_000]:
    for coherent in [True, False]:
        rows = cloud(n, 1, coherent)
        for grid in [(64, 64, 64), (128, 128, 128)]:
            V = grid[0] * grid[1] * grid[2]
            dt = t(lambda: voxel_ops.voxelize_clouds(rows[:, :3], None, grid, rows[:, 3], [15], want=("occ", "occ_keep"), occ_dtype=torch.float32))
            bytes_alg = 56 * n + 24 * V
            print(f"N={n:>9d} {'coherent' if coherent else 'shuffled':9s} grid={grid[0]:3d}^3: {dt*1e6:9.1f} us  {n/dt/1e6:9.1f} Mpts/s  {bytes_alg/dt/1e9:7.1f} GB/s algorithmic ({bytes_alg/dt/1e9/6551.7*100:.1f}% of HBM)")
# batched: 32 clouds of 60k points (one training batch)
rows = torch.cat([cloud(60_000, s) for s in range(32)])
off = torch.arange(0, 33, device=dev, dtype=torch.int64) * 60_000
dt = t(lambda: voxel_ops.voxelize_clouds(rows[:, :3], off, (64, 64, 64), rows[:, 3], [15], want=("occ", "occ_keep"), occ_dtype=torch.float32))
print(f"batch of 32 clouds x 60k pts -> 64^3: {dt*1e6:.1f} us  {32*60000/dt/1e6:.1f} Mpts/s  {32/dt:.0f} clouds/s")
