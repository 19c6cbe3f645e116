"""one config-2 step (B=32, 64^3, (9,5,5)) repeated a few times — target of the ncu captures"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
io = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
model = bench.kat_model(dev)
pool = bench.make_pool(dev, 0, 2, io)
for i in range(4):
    x, dp = pool[i % 2]
    for p in model.parameters():
        p.grad = None
    pred = model(x)
    pred.backward(dp)
torch.cuda.synchronize()
print("ok", [float(p.grad) for p in model.parameters() if p.grad is not None][:3])
