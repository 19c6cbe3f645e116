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

print("---- which part of the G0 pipeline needs float64? (7x7x7, worst-conditioned golden case)")
def run2(ks, x, dpred, last, p_mode, d_mode, prod_mode):
    m = mo.OracleSceneNet(mo.KAT_GENEO_NUM, ks, mo.KAT_PARAMS, mo.KAT_LAMBDAS, last)
    Ks = m.kernels(); lam = [m.lambda_eff(n).detach().double() for n in m.geneos]
    Kstar = sum(l * k for l, k in zip(lam, Ks.detach()))
    s = F.conv3d(x.float(), Kstar.float().view(1, 1, *ks), padding="same")          # f32 stencil
    p = torch.relu(torch.tanh(s)).double() if p_mode == "tanhf" else torch.relu(torch.tanh(s.double()))
    d = dpred.float().double() if d_mode == "d32" else dpred
    if prod_mode == "f32":
        G0 = (d.float() * (1 - p.float() * p.float()) * (p > 0)).double()
    else:
        G0 = (d * (1 - p * p) * (p > 0)).float().double()                          # f64 product, rounded once
    Kd = torch.zeros(1, 1, *ks, dtype=torch.float64, requires_grad=True)
    F.conv3d(x, Kd, padding="same").backward(G0)
    W = Kd.grad[0, 0]
    m.zero_grad(); Ks2 = m.kernels()
    L = sum(m.lambda_eff(n) * (Ks2[i, 0] * W).sum() for i, n in enumerate(m.geneos)); L.backward()
    return m.grads()
tag, ks = "syn32_7x7x7", (7, 7, 7)
x, _ = mo.synthetic_grids(2, (32, 32, 32), seed=1234)
dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
names = [str(s) for s in gold[f"{tag}|grads_names"]]; ref = dict(zip(names, gold[f"{tag}|grads"]))
for p_mode in ["tanhf", "tanh64"]:
    for d_mode in ["d32", "d64"]:
        for prod_mode in ["f32", "f64"]:
            g = run2(ks, x, dpred, "lambda_neg_0", p_mode, d_mode, prod_mode)
            errs = {n: abs(g[n] - ref[n]) / abs(ref[n]) for n in names if not np.isnan(ref[n])}
            w = max(errs, key=errs.get)
            print(f"p={p_mode:7s} dpred={d_mode} product={prod_mode}: worst {errs[w]:.2e} ({w.split('.')[-3]}.{w.split('.')[-1]})")
