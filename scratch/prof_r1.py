"""one eager SCENE-Net step (config 2, float64 boundary) repeated a few times: every kernel of the step shows up for ncu"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
model = bench.kat_model(dev)
pool = bench.make_pool(dev, 0, 3, torch.float64)
for i in range(4):
    x, dp = pool[i % 3]
    for p in model.parameters(): p.grad = None
    pred = model(x)
    pred.backward(dp)
torch.cuda.synchronize()
print(float(pred.sum()))
