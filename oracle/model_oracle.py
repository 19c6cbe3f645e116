"""TEST INFRASTRUCTURE ONLY — CPU (torch-on-CPU, float64 convolution) restatement of the
SCENE-Net model path: GENEO kernel synthesis -> conv3d -> observer -> criterion -> autograd.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product path (scene-net_b200/) never does.

The reference is Python/PyTorch and cannot travel to the GPU box, so this file restates it
with the same arithmetic engine the reference uses on a CPU host (ATen float32 elementwise
ops for the synthesis, ATen float64 `conv3d`, autograd for the backward).  Each function
cites the reference lines it follows (paths relative to the reference root).

PINNED: `oracle/make_golden.py` imports the real reference in the build container, runs it
and this restatement on the same inputs and stores the reference's outputs under
`tests/golden/`; `tests/test_oracle_model.py` holds the oracle to those fixtures
(kernels <= 1e-6*max|K|, pred/loss <= 1e-6, grads <= 5e-6 relative (the reference's own fp32 autograd noise is ~2e-6) — tighter than the 1e-5 bar the
CUDA path is held to) and, when /root/reference is present, to the live reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

EPS_V2 = 1e-8  # cylinder.py:151, arrow.py:215, neg_sphere.py:166

# hist_estimation.pickle (core/criterions/hist_estimation.pickle): the 10-bin histogram the
# reference's WeightedMSE loads (w_mse.py:58-60).  These are data, not code.
HIST_FREQS = [52648, 52727, 52553, 52392, 52366, 52380, 52501, 51922, 52499, 52300]
HIST_RANGES = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]

KINDS_V2 = {"cy": "cylinderv2", "cone": "arrow", "neg": "negSpherev2"}
KINDS_V1 = {"cy": "cylinder_kernel", "cone": "cone_kernel", "neg": "neg_sphere_kernel"}
PARAM_NAMES = {  # alphabetical = nn.ParameterDict order
    "cy": ["radius", "sigma"],
    "cone": ["apex", "cone_inc", "cone_radius", "radius", "sigma"],
    "neg": ["neg_factor", "radius", "sigma"],
}


# --------------------------------------------------------------------------------------
# index plumbing (the reference's meshgrid(...).T.reshape / .view dance)
# --------------------------------------------------------------------------------------
# dtype of the tap coordinates: float32 like the reference; tests set float64 (together with float64 parameters) to get the
# exact value of the same formulas and measure the reference's own float32 rounding noise against it
COORD_DTYPE = torch.float32


class exact_arithmetic:
    """with exact_arithmetic(): SYNTH[...] on float64 parameters evaluates the reference's formulas in float64"""

    def __enter__(self):
        global COORD_DTYPE
        self._saved, COORD_DTYPE = COORD_DTYPE, torch.float64

    def __exit__(self, *exc):
        global COORD_DTYPE
        COORD_DTYPE = self._saved
        return False


def _plane_coords(kx, ky):
    """plane[p, q] is evaluated at (i, j) = ((p*ky+q) % kx, (p*ky+q) // kx)
    (cylinder.py:164-171: rows of `.T.reshape(-1,2)` are n = j*kx + i, then `.view(kx,ky)`)."""
    n = torch.arange(kx * ky)
    i = (n % kx).to(COORD_DTYPE)
    j = (n // kx).to(COORD_DTYPE)
    return i, j


def _volume_coords(kz, kx, ky):
    """K.flatten()[r] is evaluated at (iz, ix, iy) = (r % kz, (r // kz) % kx, r // (kz*kx))
    (neg_sphere.py:187-197)."""
    r = torch.arange(kz * kx * ky)
    iz = (r % kz).to(COORD_DTYPE)
    ix = ((r // kz) % kx).to(COORD_DTYPE)
    iy = (r // (kz * kx)).to(COORD_DTYPE)
    return iz, ix, iy


def _d2_plane(kx, ky):
    i, j = _plane_coords(kx, ky)
    ci, cj = (kx - 1) / 2, (ky - 1) / 2
    nrm = torch.sqrt((i - ci) ** 2 + (j - cj) ** 2)  # linalg.norm(x_c, dim=1)
    return nrm ** 2


def _d2_volume(kz, kx, ky):
    iz, ix, iy = _volume_coords(kz, kx, ky)
    nrm = torch.sqrt((iz - (kz - 1) / 2) ** 2 + (ix - (kx - 1) / 2) ** 2 + (iy - (ky - 1) / 2) ** 2)
    return nrm ** 2


def _zero_sum(v, n):
    return v - torch.sum(v) / n


# --------------------------------------------------------------------------------------
# kernel synthesis (float32, differentiable)
# --------------------------------------------------------------------------------------
def cylinder_v2(p, ks):
    """cylinderv2 (cylinder.py:146-176): sigma*exp(-d^4/(2(r+eps)^2)), zero-summed, tiled over z."""
    kz, kx, ky = ks
    d2 = _d2_plane(kx, ky)
    f = p["sigma"] * torch.exp((d2 ** 2) * (-1 / (2 * (p["radius"] + EPS_V2) ** 2)))
    f = _zero_sum(f, kx * ky).view(kx, ky)
    return f.unsqueeze(0).repeat(kz, 1, 1)


def cylinder_v1(p, ks):
    """cylinder_kernel (cylinder.py:72-103): exp(-(d^2-r^2)^2/(2 sigma^2))."""
    kz, kx, ky = ks
    d2 = _d2_plane(kx, ky)
    f = torch.exp(((d2 - p["radius"] ** 2) ** 2) * (-1 / (2 * p["sigma"] ** 2)))
    f = _zero_sum(f, kx * ky).view(kx, ky)
    return f.unsqueeze(0).repeat(kz, 1, 1)


def arrow_v2(p, ks):
    """arrow (arrow.py:208-252): z < kz-hc -> cone slice h=z with rad_h = cone_radius*h*tan(clamp(inc)*pi);
    z >= kz-hc -> cylinder plane.  hc = int(apex)."""
    kz, kx, ky = ks
    d4 = _d2_plane(kx, ky) ** 2
    hc = int(p["apex"].to(torch.int).item())
    ch = kz - hc

    def plane(rad):
        f = p["sigma"] * torch.exp(d4 * (-1 / (2 * (rad + EPS_V2) ** 2)))
        return _zero_sum(f, kx * ky).view(1, kx, ky)

    inc = torch.clamp(p["cone_inc"], 0, 0.499)
    slices = [plane(p["cone_radius"] * h * torch.tan(inc * torch.pi)) for h in range(ch)]
    slices += [plane(p["radius"])] * hc
    return torch.cat(slices, dim=0)


def cone_v1(p, ks):
    """cone_kernel (arrow.py:170-205): exp(-(d^2-r^2)^2/(2 sig^2)); cylinder planes use sig=sigma,
    cone slice z=j uses sig_h = cone_radius*sin(cone_inc*pi/(2+h)) with h = ch-1-j."""
    kz, kx, ky = ks
    d2 = _d2_plane(kx, ky)
    hc = int(p["apex"].to(torch.int).item())
    ch = kz - hc

    def plane(sig):
        f = torch.exp(((d2 - p["radius"] ** 2) ** 2) * (-1 / (2 * sig ** 2)))
        return _zero_sum(f, kx * ky).view(1, kx, ky)

    slices = []
    for j in range(ch):
        h = ch - 1 - j
        slices.append(plane(p["cone_radius"] * torch.sin(p["cone_inc"] * torch.pi / (2 + h))))
    slices += [plane(p["sigma"])] * hc
    return torch.cat(slices, dim=0)


def neg_sphere_v2(p, ks):
    """negSpherev2 (neg_sphere.py:160-199): K = -nf*sigma*exp(-d^4/(2(r+eps)^2)); K -= (sum K + nf)/T."""
    kz, kx, ky = ks
    d2 = _d2_volume(kz, kx, ky)
    g = p["sigma"] * torch.exp((d2 ** 2) * (-1 / (2 * (p["radius"] + EPS_V2) ** 2)))
    k = (-p["neg_factor"]) * g
    k = k - (torch.sum(k) + p["neg_factor"]) / (kz * kx * ky)
    return k.view(kz, kx, ky)


def neg_sphere_v1(p, ks):
    """neg_sphere_kernel (neg_sphere.py:129-158): exp(-(d^2-r^2)^2/(2 sigma^2)) - mean - nf."""
    kz, kx, ky = ks
    d2 = _d2_volume(kz, kx, ky)
    g = torch.exp(((d2 - p["radius"] ** 2) ** 2) * (-1 / (2 * p["sigma"] ** 2)))
    k = _zero_sum(g, kz * kx * ky) - p["neg_factor"]
    return k.view(kz, kx, ky)


SYNTH = {
    "cylinderv2": cylinder_v2, "cylinder_kernel": cylinder_v1,
    "arrow": arrow_v2, "cone_kernel": cone_v1,
    "negSpherev2": neg_sphere_v2, "neg_sphere_kernel": neg_sphere_v1,
}


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------
class OracleSceneNet:
    """Functional restatement of SceneNet / SCENE_Net (SCENE_Net.py:121-339).

    geneos:  OrderedDict name -> (kind_class_name, {param: 0-dim float32 tensor})
             in channel order cy_*, cone_*, neg_* (SCENE_Net.py:264-275)
    lambdas: OrderedDict 'lambda_<name>' -> 0-dim float32 tensor, in lambdas_dict (alphabetical) order
    """

    def __init__(self, geneo_num, kernel_size, params: dict, lambdas: dict, last_lambda: str, v1=False):
        kinds = KINDS_V1 if v1 else KINDS_V2
        self.kernel_size = tuple(kernel_size)
        self.geneos = OrderedDict()
        for key in geneo_num:
            for i in range(geneo_num[key]):
                name = f"{key}_{i}"
                ps = OrderedDict()
                for pn in PARAM_NAMES[key]:
                    t = torch.tensor(float(params[f"{name}.{pn}"]), dtype=torch.float32)
                    t.requires_grad_(pn != "apex")
                    ps[pn] = t
                self.geneos[name] = (kinds[key], ps)
        self.last_lambda = last_lambda
        self.lambdas = OrderedDict()
        for ln in sorted(lambdas):  # nn.ParameterDict built from a dict -> sorted keys
            t = torch.tensor(float(lambdas[ln]), dtype=torch.float32)
            t.requires_grad_(ln != last_lambda)
            self.lambdas[ln] = t

    def kernels(self):
        """[G,1,kz,kx,ky] float64 (GENEO_Layer.compute_kernel, SCENE_Net.py:103-106)."""
        ks = [SYNTH[kind](ps, self.kernel_size).to(torch.float64).view(1, *self.kernel_size)
              for kind, ps in self.geneos.values()]
        return torch.stack(ks)

    def lambda_eff(self, name):
        """SCENE_Net.py:331 — evaluated in float32, left to right over lambdas_dict.values()."""
        ln = f"lambda_{name}"
        if ln == self.last_lambda:
            return 1 - sum(self.lambdas.values()) + self.lambdas[ln]
        return self.lambdas[ln]

    def forward(self, x):
        """SceneNet.forward (SCENE_Net.py:322-339); x [B,1,Z,X,Y] float64."""
        conv = F.conv3d(x, self.kernels(), padding="same")
        s = torch.zeros_like(x)
        for i, name in enumerate(self.geneos):
            s = s + self.lambda_eff(name) * conv[:, [i]]
        return torch.relu(torch.tanh(s))

    def named_trainable(self):
        out = OrderedDict()
        for name, (_, ps) in self.geneos.items():
            for pn, t in ps.items():
                out[f"geneos.{name}.geneo_params.{pn}"] = t
        for ln, t in self.lambdas.items():
            out[f"lambdas_dict.{ln}"] = t
        return out

    def grads(self):
        return OrderedDict((k, (None if t.grad is None else float(t.grad))) for k, t in self.named_trainable().items())

    def zero_grad(self):
        for t in self.named_trainable().values():
            t.grad = None


# --------------------------------------------------------------------------------------
# criterion: GENEO_Tversky_Loss (geneo_loss.py:145-161) = WeightedMSE (w_mse.py:114-151)
#            + FocalTverskyLoss (tversky_loss.py:81-95) + penalties (geneo_loss.py:36-71)
# --------------------------------------------------------------------------------------
def weight_target(y, weight_alpha=1.0, weight_epsilon=0.1):
    freqs = torch.tensor(HIST_FREQS, dtype=torch.int64, device=y.device)
    ranges = torch.tensor(HIST_RANGES, dtype=torch.float32, device=y.device)
    idx = torch.abs(torch.unsqueeze(y, -1) - ranges).argmin(dim=-1)
    hist = freqs[idx]
    fmin, fmax = freqs.min(), freqs.max()
    dens = (hist - fmin) / (fmax - fmin)
    w = torch.max(1 - weight_alpha * dens, torch.full_like(dens, weight_epsilon))
    return w / torch.mean(w)


def geneo_tversky_criterion(pred, y, lambdas, last, geneo_params, weight_alpha=1.0, weight_epsilon=0.1,
                            mse_weight=1.0, convex_weight=5.0, tversky_alpha=2.0, tversky_beta=1.0, focal_gamma=4.0,
                            tversky_smooth=1e-6):
    """lambdas: mapping name -> 0-dim tensor in lambdas_dict order; last: name of the frozen one;
    geneo_params: iterable of 0-dim tensors.  Works on any device (tests run it on CUDA tensors to
    drive the CUDA model exactly like the reference's unchanged criterion would)."""
    w = weight_target(y, weight_alpha, weight_epsilon)
    dense = torch.mean(mse_weight * w * (y - pred) ** 2)
    tp = (pred * y).sum()
    fp = ((1 - y) * pred).sum()
    fn = (y * (1 - pred)).sum()
    tv = (tp + tversky_smooth) / (tp + tversky_alpha * fp + tversky_beta * fn + tversky_smooth)
    focal = (1 - tv) ** focal_gamma
    cvx = convex_weight * (sum(torch.relu(-v) for k, v in lambdas.items() if k != last)
                           + torch.relu(-(1 - sum(lambdas.values()) + lambdas[last])))
    pos = convex_weight * sum(torch.relu(-t) for t in geneo_params)
    return dense + focal + cvx + pos


def geneo_tversky_loss(pred, y, model: "OracleSceneNet", **kw):
    return geneo_tversky_criterion(pred, y, model.lambdas, model.last_lambda,
                                   [t for _, ps in model.geneos.values() for t in ps.values()], **kw)


# --------------------------------------------------------------------------------------
# canonical inputs: the KAT parameter vector and config-2 synthetic grids
# --------------------------------------------------------------------------------------
KAT_GENEO_NUM = {"cy": 1, "cone": 1, "neg": 1}
KAT_KERNEL = (9, 5, 5)
KAT_PARAMS = {
    "cy_0.radius": 2.5, "cy_0.sigma": 1.8,
    "cone_0.apex": 4.0, "cone_0.cone_inc": 0.3, "cone_0.cone_radius": 2.0, "cone_0.radius": 3.0, "cone_0.sigma": 1.4,
    "neg_0.neg_factor": 0.2, "neg_0.radius": 8.0, "neg_0.sigma": 0.8,
}
KAT_LAMBDAS = {"lambda_cone_0": 0.3, "lambda_cy_0": 0.45, "lambda_neg_0": 0.25}
KAT_LAST = "lambda_cy_0"


def kat_model(kernel_size=KAT_KERNEL, v1=False):
    return OracleSceneNet(KAT_GENEO_NUM, kernel_size, KAT_PARAMS, KAT_LAMBDAS, KAT_LAST, v1=v1)


def synthetic_grids(batch, grid=(64, 64, 64), seed=1234, p_occ=0.016, p_gt=3e-4, dtype=torch.float64):
    """Config 2: Bernoulli occupancy / target grids from a CPU generator."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand((batch, 1, *grid), generator=g) < p_occ).to(dtype)
    y = (torch.rand((batch, 1, *grid), generator=g) < p_gt).to(dtype)
    return x, y


def fwd_bwd(model: OracleSceneNet, x, y=None, dpred=None):
    """One reference-semantics step on the CPU.  Either the criterion (y) or a fixed upstream
    gradient (dpred) drives the backward.  Returns (pred, loss_or_None, grads)."""
    model.zero_grad()
    pred = model.forward(x)
    loss = None
    if y is not None:
        loss = geneo_tversky_loss(pred, y, model)
        loss.backward()
    else:
        pred.backward(dpred)
    return pred.detach(), (None if loss is None else float(loss)), model.grads()
