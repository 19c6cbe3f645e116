import os, sys, torch
sys.path.insert(0, os.getcwd())
sys.argv = ["x"]
src = open("scratch/vox_bench.py").read().split("for n in [100_000")[0]
exec(src)
rows = cloud(1_000_000, 1, True)
for _ in range(3):
    out = voxel_ops.voxelize_clouds(rows[:, :3], None, (64, 64, 64), rows[:, 3], [15], want=("occ", "occ_keep"))
torch.cuda.synchronize()
rows = torch.cat([cloud(60_000, s, True) for s in range(32)])
off = torch.arange(0, 33, device=rows.device, dtype=torch.int64) * 60_000
for _ in range(3):
    out = voxel_ops.voxelize_clouds(rows[:, :3], off, (64, 64, 64), rows[:, 3], [15], want=("occ", "occ_keep"))
torch.cuda.synchronize()
print("ok")
