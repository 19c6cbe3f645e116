"""mask-driven occupancy forward: occupancy bits of sn_grid_prepare, parity with the dense stencil, timing"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
def t(fn, reps=20):
    for i in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
def bits_of(state, n):
    w = state[8:].view(torch.int32)[: (n + 31) // 32].cpu().numpy().astype("uint32")
    import numpy as np
    return np.unpackbits(w.view("uint8"), bitorder="little")[:n]
bad = 0
g = torch.Generator(device=dev).manual_seed(3)
shapes = [((2, 1, 7, 9, 13), (3, 3, 3)), ((1, 1, 16, 16, 16), (5, 5, 5)), ((3, 1, 20, 33, 70), (9, 5, 5)), ((2, 1, 64, 64, 64), (9, 5, 5)),
          ((2, 1, 24, 40, 128), (9, 7, 7)), ((1, 1, 33, 31, 100), (4, 6, 5)), ((1, 1, 40, 40, 40), (9, 9, 9)), ((1, 1, 32, 32, 32), (7, 15, 15)),
          ((2, 1, 9, 17, 31), (6, 5, 5)), ((1, 1, 64, 64, 256), (9, 5, 5))]
for shape, ks in shapes:
    for dens in (0.0, 0.02, 0.3):
        for dt in (torch.float64, torch.float32, torch.uint8):
            if dt == torch.uint8:
                x = (torch.rand(shape, generator=g, device=dev) < dens).to(torch.uint8)
            else:
                x = ((torch.rand(shape, generator=g, device=dev) < dens) * torch.rand(shape, generator=g, device=dev)).to(dt)
            x32, st = ops.prepare(x)
            n = x.numel()
            ref = (x != 0).flatten().cpu().numpy().astype("uint8")
            got = bits_of(st, n)
            if int(st[0]) != int(ref.sum()) or (got != ref).any():
                bad += 1; print("MASK MISMATCH", shape, dens, dt, int(st[0]), int(ref.sum()), int((got != ref).sum()))
            if not torch.equal(x32, x.float()):
                bad += 1; print("X32 MISMATCH", shape, dens, dt)
            K = torch.randn(ks, generator=g, device=dev) * 0.2
            for odt in (torch.float64, torch.float32):
                pd = ops.scenenet_fwd(x32, K, odt, mode=1)
                try:
                    pm = ops.scenenet_fwd(x32, K, odt, nnz=st, mode=2)
                    ps = ops.scenenet_fwd(x32, K, odt, mode=2)
                except Exception as ex:
                    if dens == 0.0 and dt == torch.float64 and odt == torch.float64: print("unsupported", shape, ks, ex)
                    continue
                e1 = float((pd - pm).abs().max()); e2 = float((pd - ps).abs().max())
                if not (e1 < 3e-6 and e2 < 3e-6):
                    bad += 1; print("FWD MISMATCH", shape, ks, dens, dt, odt, e1, e2)
# float64 tanh of the occupancy-driven forward against torch (identity kernel: s = x)
xs_ = (torch.rand((2, 1, 16, 16, 64), generator=g, device=dev) * torch.tensor([1e-4, 1e-2, 1.0, 12.0, 25.0, 0.3, 3.0, 0.05], device=dev).repeat(8)).float()
Kid = torch.zeros((3, 3, 3), device=dev); Kid[1, 1, 1] = 1.0
x32_, st_ = ops.prepare(xs_)
pt = ops.scenenet_fwd(x32_, Kid, torch.float64, nnz=st_, mode=2)
err = (pt - torch.tanh(xs_.double())).abs().max().item()
print("tanh64 max abs err", err, flush=True)
if not err < 1e-11: bad += 1
print("mismatches:", bad, flush=True)
for (B, grid, ks, dens) in [(32, (64,64,64), (9,5,5), d) for d in (0.0, 0.016, 0.03, 0.05, 0.1)] + [(32, (64,64,64), (9,7,7), 0.016), (8, (128,128,128), (9,9,9), 0.016), (8, (64,64,256), (9,5,5), 0.016)]:
    xs = [(torch.rand((B, 1, *grid), generator=g, device=dev) < dens).double() for _ in range(3)]
    prep = [ops.prepare(x) for x in xs]
    K = torch.randn(ks, generator=g, device=dev) * 0.1
    i = [0]
    def runf(mode, dt, use):
        i[0] += 1
        x32, st = prep[i[0] % 3]
        return ops.scenenet_fwd(x32, K, dt, nnz=st if use else None, mode=mode)
    r = {}
    for dt in (torch.float64, torch.float32):
        r[dt] = (t(lambda: runf(1, dt, False)), t(lambda: runf(2, dt, False)), t(lambda: runf(2, dt, True)))
    tp = t(lambda: ops.prepare(xs[i[0] % 3]))
    print(f"B={B:2d} grid={grid} k={ks} occ={dens:5.3f}: f64 out dense {r[torch.float64][0]:7.1f} scan {r[torch.float64][1]:7.1f} mask {r[torch.float64][2]:7.1f} | f32 out dense {r[torch.float32][0]:7.1f} scan {r[torch.float32][1]:7.1f} mask {r[torch.float32][2]:7.1f} | prepare f64 {tp:6.1f} us", flush=True)
