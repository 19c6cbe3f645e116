"""randomised shapes / kernel sizes / occupancies / dtypes: forward (dense, occupancy-driven) and tap gradient (dense,
occupancy-driven, device-selected) against float64 torch references"""
import os, sys, random, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
from scenenet_b200 import ops
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
rnd = random.Random(int(os.environ.get("SEED", "0")))
def pads(ks):
    pl = [(k - 1) // 2 for k in ks]; pr = [k - 1 - l for k, l in zip(ks, pl)]
    return (pl[2], pr[2], pl[1], pr[1], pl[0], pr[0])
bad = 0; n = int(os.environ.get("N", "150"))
for it in range(n):
    B = rnd.choice([1, 1, 2, 3]); Z = rnd.randint(1, 40); X = rnd.randint(1, 40); Y = rnd.choice([rnd.randint(1, 80), 4 * rnd.randint(1, 40)])
    ks = (rnd.randint(1, 11), rnd.randint(1, 9), rnd.randint(1, 9))
    dens = rnd.choice([0.0, 0.003, 0.016, 0.05, 0.2, 0.7, 1.0])
    g = torch.Generator(device=dev).manual_seed(it)
    x = ((torch.rand((B, 1, Z, X, Y), generator=g, device=dev) < dens) * (torch.rand((B, 1, Z, X, Y), generator=g, device=dev) + 0.5)).float()
    K = torch.randn(ks, generator=g, device=dev) * 0.2
    g0 = torch.randn(x.shape, generator=g, device=dev)
    xp = F.pad(x.double(), pads(ks))
    s_ref = F.conv3d(xp, K.double()[None, None]); p_ref = torch.relu(torch.tanh(s_ref))
    s_abs = F.conv3d(xp.abs(), K.double().abs()[None, None])
    W_ref = F.conv3d(xp[:, 0][None], g0[:, 0].double()[None])[0, 0]
    W_abs = F.conv3d(xp[:, 0].abs()[None], g0[:, 0].double().abs()[None])[0, 0]
    tag = f"it={it} B={B} grid=({Z},{X},{Y}) k={ks} occ={dens}"
    try:
        for od in (torch.float64, torch.float32):
            outs = {"dense": ops.scenenet_fwd(x, K, od, mode=1)}
            try:
                outs["sparse"] = ops.scenenet_fwd(x, K, od, mode=2)
            except Exception as e:
                if "UNSUPPORTED" not in str(e): raise
            # grid state from a float64 / float32 hand-over in turn (count + occupancy bits): device-selected kernel and
            # the mask-driven kernel forced at any occupancy
            x32, nnz = ops.prepare(x.double() if it % 2 else x)
            if int(nnz[0]) != int((x != 0).sum()):
                bad += 1; print("COUNT MISMATCH", tag, flush=True)
            outs["auto"] = ops.scenenet_fwd(x32, K, od, nnz=nnz)
            try:
                outs["mask"] = ops.scenenet_fwd(x32, K, od, nnz=nnz, mode=2)
            except Exception as e:
                if "UNSUPPORTED" not in str(e): raise
            for name, p in outs.items():
                err = (p.double() - p_ref).abs(); tol = 4e-6 * s_abs + 2e-7
                if not bool((err <= tol).all()):
                    bad += 1; print("FWD MISMATCH", tag, name, od, float(err.max()), flush=True)
        Ws = {"dense": ops.tapgrad(x, g0, ks, mode=1), "sparse": ops.tapgrad(x, g0, ks, mode=2)}
        x32, nnz = ops.prepare(x)
        Ws["auto"] = ops.tapgrad(x32, g0, ks, nnz=nnz)
        Ws["auto2"] = ops.tapgrad(x32, g0, ks, nnz=nnz)
        for name, W in Ws.items():
            err = (W - W_ref).abs(); tol = 3e-6 * W_abs + 1e-10
            if not bool((err <= tol).all()):
                bad += 1; print("TAPGRAD MISMATCH", tag, name, float(err.max()), float(tol.max()), flush=True)
    except Exception as e:
        bad += 1; print("EXCEPTION", tag, type(e).__name__, str(e)[:200], flush=True)
torch.cuda.synchronize()
print(f"fuzz done: {n} cases, {bad} problems")
