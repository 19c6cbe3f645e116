"""CPU emulation: which float32 step costs how much gradient accuracy? (reference = golden fp64 run)"""
import os, sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
from oracle import model_oracle as mo
torch.set_num_threads(8)
gold = np.load("tests/golden/ref_model.npz")

def run(ks, x, dpred, last, s_mode, g_mode):
    m = mo.OracleSceneNet(mo.KAT_GENEO_NUM, ks, mo.KAT_PARAMS, mo.KAT_LAMBDAS, last)
    Ks = m.kernels()            # [G,1,kz,kx,ky] f64 (values are f32)
    lam = [m.lambda_eff(n).detach().double() for n in m.geneos]
    Kstar = sum(l * k for l, k in zip(lam, Ks.detach()))  # f64 [1,kz,kx,ky]
    if s_mode == "f64":
        s = F.conv3d(x, Kstar.view(1, 1, *ks), padding="same")
    elif s_mode == "k32_acc64":
        s = F.conv3d(x, Kstar.float().double().view(1, 1, *ks), padding="same")
    elif s_mode == "f32":
        s = F.conv3d(x.float(), Kstar.float().view(1, 1, *ks), padding="same").double()
    if g_mode == "f64":
        p = torch.relu(torch.tanh(s))
        G0 = dpred * (1 - p ** 2) * (p > 0)
    else:
        p = torch.relu(torch.tanh(s.float()))
        G0 = (dpred.float() * (1 - p * p) * (p > 0)).double()
    # W exact in f64
    Kd = torch.zeros(1, 1, *ks, dtype=torch.float64, requires_grad=True)
    F.conv3d(x, Kd, padding="same").backward(G0)
    W = Kd.grad[0, 0]
    # param grads via autograd of the oracle synthesis: L = sum_g lam_g <K_g, W>
    m.zero_grad()
    Ks2 = m.kernels()
    L = sum(m.lambda_eff(n) * (Ks2[i, 0] * W).sum() for i, n in enumerate(m.geneos))
    L.backward()
    return m.grads(), W

for tag, ks in [("syn32_7x7x7", (7, 7, 7)), ("syn32_9x7x7", (9, 7, 7))]:
    x, _ = mo.synthetic_grids(2, (32, 32, 32), seed=1234)
    dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
    names = [str(s) for s in gold[f"{tag}|grads_names"]]
    ref = dict(zip(names, gold[f"{tag}|grads"]))
    for s_mode, g_mode in [("f64", "f64"), ("k32_acc64", "f64"), ("f32", "f64"), ("f64", "f32"), ("f32", "f32")]:
        g, W = run(ks, x, dpred, "lambda_neg_0", s_mode, g_mode)
        errs = {n.replace("geneos.", "").replace("geneo_params.", "").replace("lambdas_dict.", ""): abs(g[n] - ref[n]) / abs(ref[n]) for n in names if not np.isnan(ref[n])}
        worst = max(errs, key=errs.get)
        print(f"{tag} s={s_mode:10s} G0={g_mode}: worst {errs[worst]:.2e} ({worst})  all: " + " ".join(f"{v:.1e}" for v in errs.values()))
