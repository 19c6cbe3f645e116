"""SCENE-Net observer modules — drop-in mirror of the reference's core/models/SCENE_Net.py
(GENEO_Layer :56-113, SCENE_Net :121-226, SceneNet :229-339, SCENENetQuantile :347-415,
SCENE_Net_Class :421-466): same constructors, same `forward(x)`, same module tree and
`state_dict` keys (`geneos.<name>.geneo_params.<p>`, `lambdas_dict.lambda_<name>`), same
accessors, same RNG draws at construction — but `forward` is three CUDA launches
(kernel synthesis, [cast,] direct 3-D stencil + observer epilogue) and `backward` two
(tap-gradient reduction, parameter Jacobian^T) instead of ~100 ATen ops + cuDNN conv3d.

The criterion and the Lightning loop of the reference run unchanged on top: `forward`
returns `pred` in the input dtype, and the accessors return the live nn.Parameters so the
penalty terms of GENEO_Loss add their own autograd contributions.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn as nn

from ... import ops
from ..._lib import KIND
from .geneos import arrow, cylinder, neg_sphere
from .geneos.GENEO_kernel_torch import GENEO_kernel_torch


_SIDE_STREAMS: dict = {}


def _side_stream(device) -> torch.cuda.Stream:
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = st
    return st


_HI_STREAMS = {}


def _hi_stream(device) -> torch.cuda.Stream:
    """a HIGH-priority stream for the one-CTA synthesis kernels: their CTA is placed as soon as a slot frees up instead of
    queueing behind the thousands of CTAs of the grid preparation that runs beside them"""
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _HI_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=-1)
        _HI_STREAMS[key] = st
    return st


def _finish_backward(ctx, W):
    """tap gradient -> parameter gradients (+ the gradient all-reduce), shared by the two autograd functions"""
    x32, pred, K, lam, snap, nnz = ctx.saved_tensors[:6]
    if getattr(ctx.sync_group, "fused_with_param_grads", False):
        # dist.PeerAllReduce: the Jacobian kernel exchanges the gradients over NVLink peer memory itself (one launch)
        return ops.param_grads(ctx.spec, snap, K, lam, W, ctx.grad_scale, peer=ctx.sync_group)
    d = ops.param_grads(ctx.spec, snap, K, lam, W, ctx.grad_scale)
    if ctx.sync_group is not None:
        # the whole gradient payload is ONE flat float32 tensor: one collective right behind the Jacobian kernel,
        # no pack / unpack kernels; the parameter .grads are views into it
        if callable(ctx.sync_group):
            ctx.sync_group(d)  # a callable exchange on the flat payload
        else:
            import torch.distributed as dist
            dist.all_reduce(d, op=dist.ReduceOp.SUM, group=None if ctx.sync_group is True else ctx.sync_group)
    return d


class _ObserverLossFunction(torch.autograd.Function):
    """(loss, pred) with loss = weighted MSE + focal Tversky of pred against y (csrc/criterion.cu) — the observer and
    the criterion as ONE autograd node: the backward turns the criterion's closed-form dL/dpred straight into
    G0 = dL/dpred (1 - pred^2) [pred > 0] (sn_criterion_bwd with out_g0), so neither dL/dpred (67 MB at config 2) nor
    the separate G0 pass (168 MB of traffic) exists.  An extension next to the drop-in path (model(x) followed by the
    criterion), with bit-identical loss and gradients."""

    @staticmethod
    def forward(ctx, x, y, spec, crit_spec, write_last, grad_scale, sync_group, path_modes, *params):
        cur = ops.current_stream_obj(x.device)
        side, hi = _side_stream(x.device), _hi_stream(x.device)
        hi.wait_stream(cur)
        K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params], write_last_lambda=write_last, stream=hi)
        x32, nnz = ops.prepare(x.detach(), stream=side)
        cur.wait_stream(side)
        cur.wait_stream(hi)
        pred = ops.scenenet_fwd(x32, Kstar, x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32, nnz,
                                mode=path_modes[0])
        loss, coef, p, t = ops.criterion_fwd(pred, y.detach(), crit_spec)
        ctx.bwd_mode = path_modes[1]
        ctx.spec, ctx.crit_spec = spec, crit_spec
        ctx.grad_scale, ctx.sync_group = grad_scale, sync_group
        ctx.save_for_backward(x32, p, K, lam, snap, nnz, t, coef)
        ctx.mark_non_differentiable(pred)
        return (loss if pred.dtype == torch.float64 else loss.to(pred.dtype)), pred

    @staticmethod
    def backward(ctx, g_loss, _g_pred):
        x32, p, K, lam, snap, nnz, t, coef = ctx.saved_tensors[:8]
        g0 = ops.criterion_bwd(p, t, coef, ctx.crit_spec, grad_out=g_loss, as_g0=True)
        W = ops.tapgrad(x32, g0, ctx.spec.kernel_size, nnz=nnz, mode=ctx.bwd_mode)
        d = _finish_backward(ctx, W)
        unused = ctx.spec.unused
        grads = [d[i] if (ctx.needs_input_grad[i + 8] and i not in unused) else None for i in range(d.numel())]
        return (None,) * 8 + tuple(grads)


class _ObserverFunction(torch.autograd.Function):
    """pred = relu(tanh(conv3d_same(x, sum_g lambda_g K_g(theta_g)))) with a hand-written backward."""

    @staticmethod
    def forward(ctx, x, spec, write_last, grad_scale, sync_group, path_modes, prepared, *params):
        # kernel synthesis (one latency-bound CTA) and grid preparation (HBM-bound) are independent: the preparation
        # runs on a side stream (a parallel branch when the step is captured in a CUDA graph)
        if prepared is not None:  # several observers on the same grids (SCENENetQuantile): one preparation pass for all
            x32, nnz = prepared
            K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params], write_last_lambda=write_last)
        else:
            cur = ops.current_stream_obj(x.device)
            side, hi = _side_stream(x.device), _hi_stream(x.device)
            hi.wait_stream(cur)
            K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params], write_last_lambda=write_last, stream=hi)
            x32, nnz = ops.prepare(x.detach(), stream=side)  # buffers belong to the current stream, the pass runs on `side`
            cur.wait_stream(side)
            cur.wait_stream(hi)
        # pred comes back in the caller's dtype; byte/bool occupancy inputs (an extension) give float32
        pred = ops.scenenet_fwd(x32, Kstar, x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32, nnz,
                                mode=path_modes[0])
        ctx.bwd_mode = path_modes[1]
        ctx.spec = spec
        ctx.grad_scale = grad_scale
        ctx.sync_group = sync_group
        ctx.save_for_backward(x32, pred, K, lam, snap, nnz)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        x32, pred, K, lam, snap, nnz = ctx.saved_tensors[:6]
        W = ops.scenenet_bwd(x32, pred, dpred, ctx.spec.kernel_size, nnz, mode=ctx.bwd_mode)
        d = _finish_backward(ctx, W)
        unused = ctx.spec.unused
        grads = [d[i] if (ctx.needs_input_grad[i + 7] and i not in unused) else None for i in range(d.numel())]
        return (None, None, None, None, None, None, None, *grads)


class _MultiObserverFunction(torch.autograd.Function):
    """preds [Q,B,1,Z,X,Y] of Q observers on the SAME grids (SCENENetQuantile, SCENE_Net.py:347-415) as one autograd node:
    one grid preparation, Q kernel syntheses, ONE forward launch for all observers (sn_scenenet_fwd_multi: the
    occupancy-driven kernel lists a tile's non-zero voxels once and scatters them with every observer's taps), and in the
    backward one tap-gradient reduction + parameter Jacobian per observer on the shared float32 grid / grid state."""

    @staticmethod
    def forward(ctx, x, specs, counts, grad_scales, sync_groups, mode, prepared, *params):
        x32, nnz = prepared if prepared is not None else ops.prepare(x.detach())
        Ks, lams, snaps, kstars = [], [], [], []
        off = 0
        for spec, n in zip(specs, counts):
            K, lam, Kstar, snap = ops.synth_fwd(spec, [p.detach() for p in params[off:off + n]], write_last_lambda=True)
            off += n
            Ks.append(K); lams.append(lam); snaps.append(snap); kstars.append(Kstar)
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
        preds = ops.scenenet_fwd_multi(x32, torch.stack(kstars), out_dtype, nnz=nnz, mode=mode[0])
        ctx.specs, ctx.counts, ctx.grad_scales, ctx.sync_groups, ctx.bwd_mode = specs, counts, grad_scales, sync_groups, mode[1]
        ctx.save_for_backward(x32, nnz, preds, *Ks, *lams, *snaps)
        return preds

    @staticmethod
    def backward(ctx, dpreds):
        Q = len(ctx.specs)
        x32, nnz, preds = ctx.saved_tensors[:3]
        Ks, lams, snaps = (ctx.saved_tensors[3 + i * Q:3 + (i + 1) * Q] for i in range(3))
        grads = []
        off = 0
        for q, (spec, n) in enumerate(zip(ctx.specs, ctx.counts)):
            W = ops.scenenet_bwd(x32, preds[q], dpreds[q], spec.kernel_size, nnz, mode=ctx.bwd_mode)
            peer = ctx.sync_groups[q] if getattr(ctx.sync_groups[q], "fused_with_param_grads", False) else None
            d = ops.param_grads(spec, snaps[q], Ks[q], lams[q], W, ctx.grad_scales[q], peer=peer)
            if peer is None and ctx.sync_groups[q] is not None:
                if callable(ctx.sync_groups[q]):
                    ctx.sync_groups[q](d)
                else:
                    import torch.distributed as dist
                    dist.all_reduce(d, op=dist.ReduceOp.SUM, group=None if ctx.sync_groups[q] is True else ctx.sync_groups[q])
            grads += [d[i] if (ctx.needs_input_grad[7 + off + i] and i not in spec.unused) else None for i in range(n)]
            off += n
        return (None,) * 7 + tuple(grads)


def _apex_int(layer) -> int:
    """int(apex) of a cone/arrow layer.  The reference reads it with .item() on every forward (a
    device sync, arrow.py:235); apex is frozen, so the host value is cached until the tensor changes."""
    p = layer.geneo_params['apex']
    key = (p.data_ptr(), p._version)
    cached = getattr(layer, '_apex_cache', None)
    if cached is None or cached[0] != key:
        cached = (key, int(p.detach().to(torch.int).item()))
        layer._apex_cache = cached
    return cached[1]


###############################################################
#                         GENEO Layer                         #
###############################################################

class GENEO_Layer(nn.Module):

    @property
    def has_apex(self) -> bool:
        return 'apex' in self.geneo_params

    def __init__(self, geneo_class: GENEO_kernel_torch, kernel_size: tuple = None, smart=False):
        super(GENEO_Layer, self).__init__()
        self.geneo_class = geneo_class
        self.init_from_config(smart)
        if kernel_size is not None:
            self.kernel_size = kernel_size

    def init_from_config(self, smart=False):
        config = self.geneo_class.geneo_smart_config() if smart else self.geneo_class.geneo_random_config()
        self.name = config['name']
        self.kernel_size = config['kernel_size']
        self.plot = config['plot']
        params = {}
        for pname, value in config['geneo_params'].items():
            t = torch.as_tensor(value).detach().clone().to(torch.float)
            params[pname] = nn.Parameter(t, requires_grad=pname not in config['non_trainable'])
        self.geneo_params = nn.ParameterDict(params)  # plain dict -> keys sorted (reference behaviour)

    def init_from_kwargs(self, kernel_size, kwargs):
        self.kernel_size = kernel_size
        self.name = 'GENEO'
        self.plot = False
        params = {p: nn.Parameter(torch.tensor(kwargs[p], dtype=torch.float)) for p in self.geneo_class.mandatory_parameters()}
        self.geneo_params = nn.ParameterDict(params)

    def compute_kernel(self) -> torch.Tensor:
        """[1,kz,kx,ky] float64, differentiable w.r.t. geneo_params (SCENE_Net.py:103-106)."""
        if self.geneo_class is neg_sphere.negSpherev2:
            geneo = self.geneo_class(self.name, self.kernel_size, **self.geneo_params)
        else:
            geneo = self.geneo_class(self.name, self.kernel_size, plot=self.plot, **self.geneo_params)
        kernel = geneo.kernel.to(dtype=torch.double)
        return kernel.view(1, *kernel.shape)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # dead code in the reference (SCENE_Net.py:108-113 reads an undefined self.device)
        raise AttributeError("'GENEO_Layer' object has no attribute 'device'")


###############################################################
#                         SCENE-Nets                          #
###############################################################

class _SceneNetBase(nn.Module):
    """Shared body of SCENE_Net (v1 kernels) and SceneNet (v2 kernels)."""

    _classes: dict = {}

    def _build(self, geneo_num, kernel_size, plot, lam_min, lam_max, lam_device=None):
        self.sizes = {'cy': 1, 'cone': 1, 'neg': 1} if geneo_num is None else geneo_num
        if kernel_size is not None:
            self.kernel_size = kernel_size
        self.geneos: Mapping[str, GENEO_Layer] = nn.ModuleDict()
        for key in self.sizes:
            if key in self._classes:
                for i in range(self.sizes[key]):
                    self.geneos[f'{key}_{i}'] = GENEO_Layer(self._classes[key], kernel_size=kernel_size)

        # --- convex coefficients: same draws, same order as SCENE_Net.py:278-293 ---
        num_lambdas = sum(self.sizes.values())
        if lam_device is not None:  # SCENE_Net draws on its device's generator (SCENE_Net.py:177)
            lambdas = (lam_max - lam_min) * torch.rand(num_lambdas, device=lam_device, dtype=torch.float) + lam_min
        else:
            lambdas = (lam_max - lam_min) * torch.rand(num_lambdas, dtype=torch.float) + lam_min
        self.lambdas = [nn.Parameter(lamb) for lamb in lambdas]
        self.lambda_names = [f'lambda_{key}_{i}' for key, val in self.sizes.items() for i in range(val)]
        self.last_lambda = self.lambda_names[torch.randint(0, num_lambdas, (1,))[0]]
        if plot:
            print(f"last cvx_coeff: {self.last_lambda}")
        lambdas_dict = dict(zip(self.lambda_names, self.lambdas))
        lambdas_dict[self.last_lambda] = nn.Parameter(
            1 - sum(lambdas_dict.values()) + lambdas_dict[self.last_lambda], requires_grad=False)
        self.lambdas_dict = nn.ParameterDict(lambdas_dict)
        #: multiply every parameter gradient by this (1/world_size gives DDP-mean semantics without a second pass)
        self.grad_scale = 1.0
        #: None = no collective (single GPU, or DDP/Lightning does it); True / a process group = all-reduce (SUM) the
        #: flat gradient payload inside backward (use with grad_scale = 1/world for the DDP mean)
        self.grad_sync_group = None
        #: (forward, backward) kernel choice: 0 = decided on the device from the grid's non-zero count (both kernels are
        #: enqueued, one returns at once), 1 = dense stencil, 2 = occupancy-driven kernel.  Both are correct at any
        #: occupancy; a step captured for replay can pin the choice (graphs.GraphedStep(specialize=True))
        self.path_modes = (0, 0)
        if plot:
            print(f"Total Number of train params = {self.get_num_total_params()}")

    # ---- accessors (SCENE_Net.py:194-207, 299-320) -------------------------------------
    def get_geneo_nums(self):
        return self.sizes

    def get_cvx_coefficients(self):
        return self.lambdas_dict

    def get_num_total_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_model_parameters(self, detach=False):
        if detach:
            return {name: param.detach().clone() for name, param in self.named_parameters()}
        return {name: param for name, param in self.named_parameters()}

    def get_geneo_params(self):
        return nn.ParameterDict(dict([(name.replace('.', '_'), p) for name, p in self.named_parameters() if 'lambda' not in name]))

    def get_dict_parameters(self):
        return dict([(n, param.data.item()) for n, param in self.named_parameters()])

    def get_model_parameters_in_dict(self):
        ddd = {}
        for key, val in self.named_parameters():
            key_split = key.split('.')
            parameter_name = f"{key_split[-3]}.{key_split[-1]}" if 'geneo' in key else key_split[-1]
            ddd[parameter_name] = val.data.item()
        return ddd

    # ---- the hot path -------------------------------------------------------------------
    def _spec_and_params(self):
        """(ObserverSpec, parameter list).  The spec only depends on the module structure, int(apex) of the cone layers and
        which lambda is last: it is cached on those (building it was ~40 us of host time per forward)."""
        names = list(self.geneos.keys())
        layers = [self.geneos[n] for n in names]
        key = (len(names), self.last_lambda, tuple(tuple(l.kernel_size) for l in layers),
               tuple(_apex_int(l) if l.has_apex else -1 for l in layers))
        cached = self.__dict__.get('_spec_cache')
        if cached is not None and cached[0] == key:
            params = [l.geneo_params[p] for l in layers for p in l.geneo_class.abi_params]
            params.extend(self.lambdas_dict[f'lambda_{n}'] for n in names)
            return cached[1], params
        spec, params = self._build_spec_and_params(names, layers)
        self.__dict__['_spec_cache'] = (key, spec)
        return spec, params

    def _build_spec_and_params(self, names, layers):
        ks = {tuple(int(v) for v in l.kernel_size) for l in layers}
        if len(ks) != 1:
            raise ValueError(f"all GENEO kernels of an observer must share one kernel_size, got {ks}")
        kinds = [KIND[l.geneo_class.kind_name] for l in layers]
        params = []
        unused = set()
        kz = next(iter(ks))[0]
        for l, kind in zip(layers, kinds):
            if kind in (KIND["cone_kernel"], KIND["arrow"]):
                hc = _apex_int(l)
                if hc < 0 or hc > kz:
                    raise ValueError(f"int(apex) = {hc} must lie in [0, kernel_size[0] = {kz}] ({l.name})")
                unused |= ops.unused_cone_params(kind, hc, kz, len(params))
            params.extend(l.geneo_params[p] for p in l.geneo_class.abi_params)
        lam_keys = list(self.lambdas_dict.keys())  # iteration order of sum(self.lambdas_dict.values())
        order = [names.index(k[len('lambda_'):]) for k in lam_keys]
        params.extend(self.lambdas_dict[f'lambda_{n}'] for n in names)
        spec = ops.ObserverSpec(kinds=kinds, kernel_size=ks.pop(), lambda_sum_order=order,
                                last_lambda=names.index(self.last_lambda[len('lambda_'):]), observer=True,
                                unused=frozenset(unused))
        return spec, params

    def forward(self, x: torch.Tensor, _prepared=None) -> torch.Tensor:
        """same signature as the reference; `_prepared` (internal) = the result of ops.prepare(x) shared between
        several observers that read the same grids"""
        spec, params = self._spec_and_params()
        if not params[0].is_cuda:
            raise RuntimeError("scenenet_b200: the model lives on the CPU; move it with .cuda() — the hot path is "
                               "CUDA-only (no CPU fallback)")
        if x.device != params[0].device:
            raise RuntimeError(f"input on {x.device} but model on {params[0].device}")
        # write_last=True reproduces the side effect of SCENE_Net.py:333 (last lambda <- 1 - sum(others)), in place
        return _ObserverFunction.apply(x, spec, True, float(self.grad_scale), self.grad_sync_group, tuple(self.path_modes),
                                       _prepared, *params)

    def forward_with_criterion(self, x: torch.Tensor, y: torch.Tensor, crit_spec: "ops.CriterionSpec"):
        """(loss, pred): forward + the fused criterion (weighted MSE + focal Tversky) as one autograd node — see
        _ObserverLossFunction.  Extension; use through `GENEO_Tversky_Loss.training_loss(model, x, y)`."""
        spec, params = self._spec_and_params()
        if not params[0].is_cuda:
            raise RuntimeError("scenenet_b200: the model lives on the CPU; move it with .cuda() — the hot path is "
                               "CUDA-only (no CPU fallback)")
        if x.device != params[0].device or y.device != x.device:
            raise RuntimeError(f"input on {x.device} / target on {y.device} but model on {params[0].device}")
        if y.shape != x.shape:
            raise ValueError(f"target shape {tuple(y.shape)} differs from the grid shape {tuple(x.shape)}")
        return _ObserverLossFunction.apply(x, y, spec, crit_spec, True, float(self.grad_scale), self.grad_sync_group,
                                           tuple(self.path_modes), *params)


class SCENE_Net(_SceneNetBase):
    """v1 kernels (cylinder_kernel / cone_kernel / neg_sphere_kernel), lambdas ~ U(0, 0.6)."""

    _classes = {'cy': cylinder.cylinder_kernel, 'cone': arrow.cone_kernel, 'neg': neg_sphere.neg_sphere_kernel}

    def __init__(self, geneo_num=None, kernel_size=None, plot=False,
                 device=torch.device('cuda' if torch.cuda.is_available() else 'cpu')):
        super(SCENE_Net, self).__init__()
        self.device = device
        self._build(geneo_num, kernel_size, plot, lam_min=0.0, lam_max=0.6, lam_device=device)


class SceneNet(_SceneNetBase):
    """v2 kernels (cylinderv2 / arrow / negSpherev2), lambdas ~ U(-2/G, 1/G)."""

    _classes = {'cy': cylinder.cylinderv2, 'cone': arrow.arrow, 'neg': neg_sphere.negSpherev2}

    def __init__(self, geneo_num=None, kernel_size=None, plot=False):
        super(SceneNet, self).__init__()
        sizes = {'cy': 1, 'cone': 1, 'neg': 1} if geneo_num is None else geneo_num
        n = sum(sizes.values())
        self._build(geneo_num, kernel_size, plot, lam_min=-2 / n, lam_max=1 / n)


###############################################################
#                     SCENE-Net Quantile                      #
###############################################################
class SCENENetQuantile(nn.Module):
    """One SCENE_Net per quantile on the same input (SCENE_Net.py:347-415)."""

    def __init__(self, geneo_num=None, kernel_size=None, qs=torch.tensor([0.1, 0.5, 0.9]), plot=False, model_path=None,
                 device=torch.device('cuda' if torch.cuda.is_available() else 'cpu')) -> None:
        super(SCENENetQuantile, self).__init__()
        if model_path is not None:
            raise NotImplementedError("legacy .pt checkpoints (SCENE_Net.py:18-49) are outside the hot path")
        self.scnets = nn.ModuleList([SCENE_Net(geneo_num, kernel_size, plot) for _ in range(len(qs))]).to(device)
        self.qs = qs
        self.device = device
        #: True: one observer after the other like the reference (SCENE_Net.py:409-415) instead of the fused forward
        self.per_observer_forward = False

    def get_num_total_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_dict_parameters(self):
        return dict([(n, param.data.item()) for n, param in self.named_parameters()])

    def get_cvx_coefficients(self):
        return [scnet.get_cvx_coefficients() for scnet in self.scnets]

    def get_geneo_params(self):
        return [scnet.get_geneo_params() for scnet in self.scnets]

    def forward(self, x: torch.Tensor):
        # SCENE_Net.py:409-415: every quantile's observer reads the same grids — the float32 copy and the grid state are
        # produced once and shared (the observers then differ only in their 13 scalars)
        prepared = ops.prepare(x.detach()) if x.is_cuda else None
        nets = list(self.scnets)
        sizes = {tuple(net._spec_and_params()[0].kernel_size) for net in nets} if x.is_cuda else set()
        if x.is_cuda and len(sizes) == 1 and 1 < len(nets) <= 8 and x.numel() and not self.per_observer_forward:
            # ONE forward for all quantiles (sn_scenenet_fwd_multi) — the occupancy-driven kernel lists the non-zero voxels
            # of a tile once and scatters them with every observer's taps; with gradients enabled the same launch is one
            # autograd node (_MultiObserverFunction) whose backward runs one tap-gradient reduction per observer on the
            # shared grid state.  Same values as the per-observer path below (same kernels, same summation order).
            specs, counts, flat = [], [], []
            for net in nets:
                spec, params = net._spec_and_params()
                specs.append(spec); counts.append(len(params)); flat.extend(params)
            modes = {tuple(net.path_modes) for net in nets}
            mode = modes.pop() if len(modes) == 1 else (0, 0)
            preds = _MultiObserverFunction.apply(x, tuple(specs), tuple(counts), tuple(float(n.grad_scale) for n in nets),
                                                 tuple(n.grad_sync_group for n in nets), mode, prepared, *flat)  # [Q,B,1,Z,X,Y]
            return preds[:, :, 0].permute(1, 0, 2, 3, 4).to(torch.float32).contiguous()
        return torch.cat([net(x, _prepared=prepared).to(torch.float32) for net in nets], dim=1)


class SCENE_Net_Class(nn.Module):
    """Thresholded classifier (SCENE_Net.py:421-466): (gnet(x) >= tau) as 0/1 in x.dtype."""

    def __init__(self, geneo_num=None, plot=True, gnet_requires_grad=True, gnet_model_path=None):
        super().__init__()
        if gnet_model_path is not None:
            raise NotImplementedError("legacy checkpoints are outside the hot path")
        self.gnet = SCENE_Net(geneo_num)
        if not gnet_requires_grad:
            for param in self.gnet.parameters():
                param.requires_grad = False
        tau_min, tau_max = 0.2, 0.6
        self.tau = nn.Parameter((tau_max - tau_min) * torch.rand(1, dtype=torch.float)[0])

    def get_threshold(self):
        return self.tau

    def get_geneo_nums(self):
        return self.gnet.sizes

    def get_cvx_coefficients(self):
        return self.gnet.lambdas_dict

    def get_geneo_params(self):
        return nn.ParameterDict(dict([(name.replace('.', '_'), p) for name, p in self.gnet.named_parameters() if 'lambda' not in name]))

    def get_dict_parameters(self):
        return dict([(n, param.data.item()) for n, param in self.gnet.named_parameters()])

    def _tau_host(self) -> float:
        """host copy of tau, refreshed only when the parameter is rewritten (an optimizer step, load_state_dict):
        `float(self.tau)` on every forward would be a device synchronisation (the reference compares on the device,
        SCENE_Net.py:465-466); same caching as `_apex_int`"""
        p = self.tau
        key = (p.data_ptr(), p._version)
        cached = self.__dict__.get('_tau_cache')
        if cached is None or cached[0] != key:
            cached = (key, float(p.detach()))
            self.__dict__['_tau_cache'] = cached
        return cached[1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.threshold(self.gnet(x), self._tau_host()).to(x.dtype)
