"""forward kernel timings (CUDA events, rotating inputs > L2): dense / occupancy-driven / per-tile auto, at several occupancies
and on rolled copies of the sample_575 grid (clustered)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0)
ks = (9, 5, 5)
g = torch.Generator(device=dev).manual_seed(1)
K = (torch.randn(ks, generator=g, device=dev) * 0.2)

def timeit(fn, n_sets, reps=30):
    for i in range(4): fn(i % n_sets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i % n_sets)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3

def bench(name, xs, odt=torch.float64):
    prep = [ops.prepare(x) for x in xs]
    n = len(xs)
    out = {}
    for mode, tag in ((1, "dense"), (2, "sparse"), (0, "auto")):
        try:
            out[tag] = round(timeit(lambda i: ops.scenenet_fwd(prep[i][0], K, odt, nnz=prep[i][1], mode=mode), n), 1)
        except Exception as e:
            out[tag] = repr(e)[:60]
    print(name, str(odt).split('.')[-1], json.dumps(out), flush=True)

import subprocess
def clocks():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                              capture_output=True, text=True).stdout.strip()
    except Exception as e:
        return repr(e)
print("clocks before:", clocks())
print("fp32 peak probe TFLOP/s:", ops.fp32_peak_probe(2000, dev), "clocks:", clocks(), flush=True)
a_ = torch.empty(1 << 28, dtype=torch.float32, device=dev); b_ = torch.empty_like(a_)
print("copy 1 GiB -> GB/s:", 2 * a_.numel() * 4 / (timeit(lambda i: b_.copy_(a_), 1, 10) * 1e-6) / 1e9, flush=True)
del a_, b_
shape = (32, 1, 64, 64, 64)
for dens in (0.0, 0.016, 0.05, 0.12):
    xs = [(torch.rand(shape, generator=g, device=dev) < dens).double() for _ in range(4)]
    bench(f"bernoulli {dens}", xs)
xs = [(torch.rand(shape, generator=g, device=dev) < 0.016).to(torch.uint8) for _ in range(4)]
bench("bernoulli 0.016 uint8 in / f32 out", xs, torch.float32)
v = np.load("tests/golden/vox_sample_575.npz")
x575 = np.zeros(64 ** 3); x575[v["restated_density_idx"]] = 1.0
base = torch.from_numpy(x575).view(64, 64, 64).to(dev)
xs = []
for s in range(4):
    xs.append(torch.stack([torch.roll(base, shifts=(0, 3 * b + s, 5 * b + 2 * s), dims=(0, 1, 2)) for b in range(32)])[:, None].contiguous())
bench("sample_575 x 32 (rolled)", xs)
# KITTI-shaped layer: 2.8 % overall, one dense layer
xs = []
for s in range(4):
    x = (torch.rand((8, 1, 256, 64, 64), generator=g, device=dev) < 0.005).double()
    x[:, :, 100:104] = (torch.rand((8, 1, 4, 64, 64), generator=g, device=dev) < 0.35).double()
    xs.append(x)
bench("kitti-like 8 x (256,64,64)", xs)
# non-binary values
xs = [((torch.rand(shape, generator=g, device=dev) < 0.016) * torch.rand(shape, generator=g, device=dev)).double() for _ in range(4)]
bench("density grids 0.016", xs)

print("fp32 peak probe TFLOP/s:", ops.fp32_peak_probe(2000, dev), "clocks after:", clocks(), flush=True)
