"""dense-x case of tests/test_gpu_model.py::test_dense_float_input_and_f32_dtype, emulated on the CPU:
float32 conv + float64 tanh/G0 rounded to float32 + EXACT (float64) tap-gradient accumulation."""
import os, sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
import warnings; warnings.filterwarnings("ignore")
from oracle import model_oracle as mo
import scenenet_b200 as sb
torch.set_num_threads(8)
geneo_num, ks, grid, B, seed = {'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), (32, 32, 32), 2, 5
torch.manual_seed(seed)
m = sb.SceneNet(dict(geneo_num), tuple(ks))
params = {f"{n}.{pn}": float(p) for n, l in m.geneos.items() for pn, p in l.geneo_params.items()}
lambdas = {k: float(v) for k, v in m.lambdas_dict.items()}
g = torch.Generator().manual_seed(seed + 100)
x = torch.rand((B, 1, *grid), generator=g, dtype=torch.float64)
dpred = torch.randn(x.shape, generator=g, dtype=torch.float64)
o = mo.OracleSceneNet(dict(geneo_num), ks, params, lambdas, m.last_lambda)
pr, _, ref = mo.fwd_bwd(o, x, None, dpred)

def emul(acc_chunk):
    o2 = mo.OracleSceneNet(dict(geneo_num), ks, params, lambdas, m.last_lambda)
    Ks = o2.kernels().detach()
    lam = [o2.lambda_eff(n).detach().double() for n in o2.geneos]
    Kstar = sum(l * k for l, k in zip(lam, Ks)).float()
    x32 = x.float()
    s = F.conv3d(x32, Kstar.view(1, 1, *ks), padding="same")
    p = torch.relu(torch.tanh(s.double()))
    G0 = (dpred * (1 - p * p) * (p > 0)).float()
    # tap gradient with controllable accumulation: products in f32, sums in chunks of `acc_chunk` voxels in f32, then f64
    xp = F.pad(x32, (2, 2, 2, 2, 4, 4))
    W = torch.zeros(ks, dtype=torch.float64)
    g0 = G0.reshape(-1)
    for dz in range(ks[0]):
        for dx in range(ks[1]):
            for dy in range(ks[2]):
                xs = xp[:, :, dz:dz + grid[0], dx:dx + grid[1], dy:dy + grid[2]].reshape(-1)
                prod = g0 * xs                      # f32 products (one rounding, like FFMA's single rounding is better)
                if acc_chunk == 0:
                    W[dz, dx, dy] = prod.double().sum()
                else:
                    n = prod.numel() // acc_chunk * acc_chunk
                    part = prod[:n].view(-1, acc_chunk)
                    # sequential f32 accumulation inside a chunk
                    a = torch.zeros(part.shape[0], dtype=torch.float32)
                    for j in range(acc_chunk):
                        a = a + part[:, j]
                    W[dz, dx, dy] = a.double().sum() + prod[n:].double().sum()
    o2.zero_grad()
    Ks2 = o2.kernels()
    L = sum(o2.lambda_eff(n) * (Ks2[i, 0] * W).sum() for i, n in enumerate(o2.geneos))
    L.backward()
    got = o2.grads()
    errs = {n: abs(got[n] - r) / abs(r) for n, r in ref.items() if r is not None}
    w = max(errs, key=errs.get)
    return errs[w], w

for chunk in [0, 32, 128, 512]:
    e, w = emul(chunk)
    print(f"acc chunk {chunk:4d}: worst grad rel err {e:.2e} ({w})")
