import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import ops
from oracle import model_oracle as mo, ref_shim
dev = "cuda"
torch.manual_seed(0)
m = sb.SceneNet(dict(mo.KAT_GENEO_NUM), (9, 5, 5)).to(dev)
ref_shim.set_scenenet_params(m, mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST)
spec, params = m._spec_and_params()
K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params])
print("type", type(Kstar), "k64 tail", Kstar.k64[-3:].tolist(), "max|K|", float(Kstar.abs().max()))
o = mo.kat_model()
Ks = o.kernels().detach().to(dev); lam64 = torch.stack([o.lambda_eff(n).detach().double() for n in o.geneos]).to(dev)
K64 = (Ks[:, 0] * lam64.view(-1, 1, 1, 1)).sum(0)
print("k64 vs oracle:", float((Kstar.k64[:-1].view(9, 5, 5) - K64).abs().max()))
# 1) exact path unit test: every non-zero sum re-evaluated (k64[T] huge)
g = torch.Generator(device=dev).manual_seed(3)
xs = (torch.rand((2, 1, 24, 24, 64), generator=g, device=dev) < 0.05).double()
x32, st = ops.prepare(xs)
want = torch.relu(torch.tanh(F.conv3d(xs, K64[None, None], padding="same")))
for mode in (1, 2):
    Kt = Kstar.clone().as_subclass(ops.KstarTensor); k = Kstar.k64.clone(); k[-1] = 1e30; Kt.k64 = k
    got = ops.scenenet_fwd(x32, Kt, torch.float64, nnz=st, mode=mode)
    plain = ops.scenenet_fwd(x32, Kstar, torch.float64, nnz=st, mode=mode)
    print(f"mode {mode}: all-exact max err {float((got - want).abs().max()):.3e}; normal max err {float((plain - want).abs().max()):.3e}")
# 2) the config-2 batch
x, _ = mo.synthetic_grids(32, (64, 64, 64), seed=1234)
xd = x.to(dev)
s_ref = (F.conv3d(xd, Ks, padding="same") * lam64.view(1, -1, 1, 1, 1)).sum(1, keepdim=True)
x32, st = ops.prepare(xd)
for mode in (0, 1, 2):
    pred = ops.scenenet_fwd(x32, Kstar, torch.float64, nnz=st, mode=mode)
    fl = ((pred > 0) != (s_ref > 0))
    print(f"mode {mode}: flips {int(fl.sum())}", [(tuple(i.tolist()), float(pred[tuple(i)]), float(s_ref[tuple(i)])) for i in fl.nonzero()[:5]])
    nk = Kstar.clone()  # plain tensor: no k64 -> no re-evaluation
    pred0 = ops.scenenet_fwd(x32, nk.as_subclass(torch.Tensor), torch.float64, nnz=st, mode=mode)
    fl0 = ((pred0 > 0) != (s_ref > 0))
    print(f"   without k64: flips {int(fl0.sum())}", [(tuple(i.tolist()), float(pred0[tuple(i)]), float(s_ref[tuple(i)])) for i in fl0.nonzero()[:5]])
for idx in [(13, 0, 8, 42, 57), (14, 0, 13, 54, 49)]:
    print(idx, "s_ref", float(s_ref[idx]))
