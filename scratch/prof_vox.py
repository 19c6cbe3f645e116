import os, sys, torch
sys.path.insert(0, os.getcwd())
sys.argv = ["x"]
import importlib.util
spec = importlib.util.spec_from_file_location("vb", "scratch/vox_bench.py")
src = open("scratch/vox_bench.py").read().split("for n in [100_000")[0]
exec(src)
rows = cloud(10_000_000, 1, True)
for _ in range(3):
    out = voxel_ops.voxelize_clouds(rows[:, :3], None, (128, 128, 128), rows[:, 3], [15], want=("occ", "occ_keep"), occ_dtype=torch.float32)
torch.cuda.synchronize()
print(int(out["count"].sum()))
