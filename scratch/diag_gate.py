"""Why do the config-2 full-size gradients differ by 2e-4 from the reference?  Hypothesis: relu-gate flips at s ~ 0."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import ops
from oracle import model_oracle as mo, ref_shim
dev = "cuda"
x, _ = mo.synthetic_grids(32, (64, 64, 64), seed=1234)
dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
torch.manual_seed(0)
m = sb.SceneNet(dict(mo.KAT_GENEO_NUM), (9, 5, 5)).to(dev)
ref_shim.set_scenenet_params(m, mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST)
xd, dd = x.to(dev), dpred.to(dev)
# float64 reference sum on the GPU from the oracle's float32 kernels (same arithmetic as the reference's conv)
o = mo.kat_model()
Ks = o.kernels().detach().to(dev)                       # [G,1,kz,kx,ky] f64
lam = torch.stack([o.lambda_eff(n).detach().double() for n in o.geneos]).to(dev)
conv = F.conv3d(xd, Ks, padding="same")                  # [B,G,Z,X,Y]
s_ref = (conv * lam.view(1, -1, 1, 1, 1)).sum(1, keepdim=True)
p_ref = torch.relu(torch.tanh(s_ref))
for modes in ((1, 1), (2, 2)):
    m.path_modes = modes
    for p in m.parameters():
        p.grad = None
    pred = m(xd)
    flips = ((pred > 0) != (p_ref > 0))
    print(f"modes {modes}: gate flips {int(flips.sum())}, |s_ref| at flips {s_ref[flips].abs().tolist()[:8]}, dpred at flips {dd[flips].tolist()[:8]}")
    print("  near-zero: #|s_ref|<1e-6:", int(((s_ref.abs() < 1e-6) & (s_ref != 0)).sum()), " #<1e-7:", int(((s_ref.abs() < 1e-7) & (s_ref != 0)).sum()),
          " #<1e-5:", int(((s_ref.abs() < 1e-5) & (s_ref != 0)).sum()))
    pred.backward(dd)
    g = {n: (None if p.grad is None else float(p.grad)) for n, p in m.named_parameters()}
    # gradients with the gate taken from the reference: G0' = dpred (1 - pred^2) [p_ref > 0]
    G0 = (dd * (1 - pred.detach() ** 2) * (p_ref > 0)).float()
    x32, nnz = ops.prepare(xd)
    W = ops.tapgrad(x32, G0, (9, 5, 5), nnz=nnz)
    spec, params = m._spec_and_params()
    K, lam_, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params])
    d = ops.param_grads(spec, snap, K, lam_, W, 1.0)
    print("  grads:", {k.split('.')[-1] + '@' + k.split('.')[1]: v for k, v in g.items() if v is not None})
    print("  grads with the reference's gate:", d.tolist())
