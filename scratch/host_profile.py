"""where does the host spend its time in an eager step? (cProfile over 300 eager fwd+bwd steps at config 2)"""
import cProfile, os, pstats, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
model = bench.kat_model(dev)
pool = bench.make_pool(dev, 0, 3, torch.float64)
def step(i):
    x, dp = pool[i % 3]
    for p in model.parameters(): p.grad = None
    pred = model(x)
    pred.backward(dp)
for i in range(20): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(300): step(i)
torch.cuda.synchronize()
print(f"eager: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us/step")
pr = cProfile.Profile(); pr.enable()
for i in range(300): step(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
