"""profiling target: the two forward kernels on the config-2 batch (KAT kernel with its float64 twin)"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import scenenet_b200 as sb
from scenenet_b200 import ops
from oracle import model_oracle as mo, ref_shim
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dens = float(os.environ.get("DENS", "0.016"))
torch.manual_seed(0)
m = sb.SceneNet(dict(mo.KAT_GENEO_NUM), (9, 5, 5)).to(dev)
ref_shim.set_scenenet_params(m, mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST)
spec, params = m._spec_and_params()
K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params])
g = torch.Generator(device=dev).manual_seed(1)
xs = [(torch.rand((32, 1, 64, 64, 64), generator=g, device=dev) < dens).double() for _ in range(3)]
prep = [ops.prepare(x) for x in xs]
for i in range(4):
    pd = ops.scenenet_fwd(prep[i % 3][0], Kstar, torch.float64, nnz=prep[i % 3][1], mode=1)
    ps = ops.scenenet_fwd(prep[i % 3][0], Kstar, torch.float64, nnz=prep[i % 3][1], mode=2)
torch.cuda.synchronize()
print(float(pd.sum()), float(ps.sum()), float((pd - ps).abs().max()))
