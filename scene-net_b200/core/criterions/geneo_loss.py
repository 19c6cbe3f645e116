"""GENEO losses — drop-in mirror of the reference's core/criterions/geneo_loss.py (GENEO_Loss :24-91,
GENEO_Tversky_Loss :145-168): same constructors and `forward(y_pred, y_gt, cvx_coeffs, geneo_params)`.

`forward` is ONE autograd node of three launches (fused reduction over (pred, y), finalisation, one penalty kernel
over the live parameters that also adds its two terms to the loss scalar) instead of ~80 ATen ops; the backward is one
elementwise kernel for dL/dpred plus one scaled copy of the penalty gradients.  The penalty terms act on the live nn.Parameters
handed in, exactly like the reference's (`cvx_loss`, `positive_regularizer`).
"""
from __future__ import annotations

import torch

from ... import ops
from .tversky_loss import FocalTverskyLoss
from .w_mse import WeightedMSE, _FusedCriterion


_ROLE_MASKS: dict = {}


def _cvx_mask(roles, device):
    """bool mask 'parameter i is a convex coefficient', cached per role pattern (no host->device copy on the
    steady-state path, so the criterion can be captured in a CUDA graph after one eager warm-up step)"""
    key = (tuple(roles), str(device))
    m = _ROLE_MASKS.get(key)
    if m is None:
        m = torch.tensor([r != 0 for r in roles], dtype=torch.bool, device=device)
        _ROLE_MASKS[key] = m
    return m


class _PenaltyFunction(torch.autograd.Function):
    """(weight * cvx_loss, weight * positive_regularizer) of the parameters, float32 (geneo_loss.py:36-71)."""

    @staticmethod
    def forward(ctx, roles, weight, *params):
        out = ops.param_penalty([p.detach() for p in params], roles, weight)
        ctx.save_for_backward(out)
        ctx.is_cvx = _cvx_mask(roles, out.device)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_cvx, g_pos):
        (out,) = ctx.saved_tensors
        d = out[2:] * torch.where(ctx.is_cvx, g_cvx.to(torch.float32), g_pos.to(torch.float32))
        grads = [d[i] if ctx.needs_input_grad[2 + i] else None for i in range(d.numel())]
        return (None, None, *grads)


def _collect(cvx_coeffs, geneo_params):
    """(parameter list, role list) in the reference's summation order (geneo_loss.py:36-71)"""
    params, roles = [], []
    if len(cvx_coeffs) > 0:
        # geneo_loss.py:48: the frozen coefficient is the one computed from the others
        last_phi = [phi_name for phi_name in cvx_coeffs if not cvx_coeffs[phi_name].requires_grad][0]
        for name, phi in cvx_coeffs.items():
            params.append(phi)
            roles.append(1 if name == last_phi else 2)
    for g in geneo_params.values():
        params.append(g)
        roles.append(0)
    return params, roles


class _FusedGeneoCriterion(torch.autograd.Function):
    """The whole GENEO loss as ONE autograd node: [weighted MSE] + [focal Tversky] + cvx penalty + positive regulariser.
    Forward = three launches (reduction over (pred, y), finalisation, penalty kernel that also adds its two terms to
    the loss scalar); backward = the closed-form dL/dpred kernel + one scaled copy of the penalty gradients.  (As
    separate nodes joined by torch adds the step carried eight more elementwise launches of autograd glue.)"""

    @staticmethod
    def forward(ctx, pred, y, spec, roles, weight, *params):
        loss, coef, p, t = ops.criterion_fwd(pred.detach(), y.detach(), spec)
        out = ops.param_penalty([q.detach() for q in params], roles, weight, loss_accum=loss)
        ctx.spec, ctx.shape = spec, pred.shape
        ctx.save_for_backward(p, t, coef, out)
        return loss if pred.dtype == torch.float64 else loss.to(pred.dtype)

    @staticmethod
    def backward(ctx, g):
        p, t, coef, out = ctx.saved_tensors
        d = None
        if ctx.needs_input_grad[0]:
            d = ops.criterion_bwd(p, t, coef, ctx.spec, grad_out=g)
            if d.shape != ctx.shape:  # pred was broadcast against y
                d = d.sum_to_size(ctx.shape)
        dp = out[2:] * g.to(torch.float32)
        grads = [dp[i] if ctx.needs_input_grad[5 + i] else None for i in range(dp.numel())]
        return (d, None, None, None, None, *grads)


def _penalties(cvx_coeffs, geneo_params, weight):
    """-> (cvx_penalty, positive_penalty) tensors; either mapping may be empty."""
    params, roles = _collect(cvx_coeffs, geneo_params)
    return _PenaltyFunction.apply(roles, float(weight), *params)


class GENEO_Loss(WeightedMSE):
    """Weighted MSE + penalties on non-positive convex coefficients / GENEO parameters."""

    def __init__(self, targets=None, weighting_scheme_path=None, weight_alpha=1, weight_epsilon=0.1, mse_weight=1, convex_weight=1, **kwargs) -> None:
        super().__init__(targets, weighting_scheme_path, weight_alpha, weight_epsilon, mse_weight, **kwargs)
        self.cvx_w = convex_weight

    def cvx_loss(self, cvx_coeffs: torch.nn.ParameterDict):
        if len(cvx_coeffs) == 0:
            return 0
        return _penalties(cvx_coeffs, {}, self.cvx_w)[0]

    def positive_regularizer(self, params: torch.nn.ParameterDict):
        if len(params) == 0:
            return 0
        return _penalties({}, params, self.cvx_w)[1]

    def _both_penalties(self, cvx_coeffs, geneo_params):
        if len(cvx_coeffs) == 0 and len(geneo_params) == 0:
            return 0, 0
        cvx, pos = _penalties(cvx_coeffs, geneo_params, self.cvx_w)
        return (cvx if len(cvx_coeffs) else 0), (pos if len(geneo_params) else 0)

    def _fused(self, y_pred, y_gt, spec, cvx_coeffs, geneo_params):
        params, roles = _collect(cvx_coeffs, geneo_params)
        if not params:  # no parameters handed in: the data terms alone
            return _FusedCriterion.apply(y_pred, y_gt, spec)
        return _FusedGeneoCriterion.apply(y_pred, y_gt, spec, roles, float(self.cvx_w), *params)

    def forward(self, y_pred: torch.Tensor, y_gt: torch.Tensor, cvx_coeffs: torch.nn.ParameterDict, geneo_params: torch.nn.ParameterDict):
        # dense_criterion + cvx_penalty + non_positive_penalty (geneo_loss.py:73-91)
        return self._fused(y_pred, y_gt, self._spec(1), cvx_coeffs, geneo_params)

    def __str__(self):
        return f"GENEO Loss with mse_weight={self.mse_weight} and alpha={self.weight_alpha} and epsilon={self.weight_epsilon}"

    @staticmethod
    def add_model_specific_args(parent_parser):
        parent_parser = WeightedMSE.add_model_specific_args(parent_parser)
        parser = parent_parser.add_argument_group('GENEO_Loss')
        parser.add_argument('--cvx_w', type=float, default=1., help='weight of the convexity penalty')
        return parent_parser


class GENEO_Tversky_Loss(GENEO_Loss):

    def __init__(self, targets=None, weighting_scheme_path=None, weight_alpha=1, weight_epsilon=0.1, mse_weight=1, convex_weight=1,
                 tversky_alpha=0.5, tversky_beta=1, focal_gamma=1, tversky_smooth=1, **kwargs) -> None:
        super().__init__(targets, weighting_scheme_path, weight_alpha, weight_epsilon, mse_weight, convex_weight, **kwargs)
        self.tversky = FocalTverskyLoss(tversky_alpha, tversky_beta, focal_gamma, tversky_smooth)

    def fused_spec(self) -> ops.CriterionSpec:
        t = self.tversky
        return self._spec(3, tversky_alpha=float(t.tversky_alpha), tversky_beta=float(t.tversky_beta),
                          focal_gamma=float(t.focal_gamma), tversky_smooth=float(t.tversky_smooth))

    def forward(self, y_pred: torch.Tensor, y_gt: torch.Tensor, cvx_coeffs: torch.nn.ParameterDict, geneo_params: torch.nn.ParameterDict):
        # dense_criterion + tversky_crit + cvx_penalty + non_positive_penalty as one node (geneo_loss.py:155-161)
        return self._fused(y_pred, y_gt, self.fused_spec(), cvx_coeffs, geneo_params)

    def training_loss(self, model, x: torch.Tensor, y_gt: torch.Tensor):
        """(loss, pred) of `model` on (x, y_gt) — extension: the observer forward and this criterion as ONE autograd
        node (SceneNet.forward_with_criterion): the backward emits G0 directly, dL/dpred is never materialised.
        Same value and gradients as `self(model(x), y_gt, model.get_cvx_coefficients(), model.get_geneo_params())`."""
        dense_and_tversky, pred = model.forward_with_criterion(x, y_gt, self.fused_spec())
        cvx_penalty, non_positive_penalty = self._both_penalties(model.get_cvx_coefficients(), model.get_geneo_params())
        return dense_and_tversky + cvx_penalty + non_positive_penalty, pred

    @staticmethod
    def add_model_specific_args(parent_parser):
        parent_parser = GENEO_Loss.add_model_specific_args(parent_parser)
        return FocalTverskyLoss.add_model_specific_args(parent_parser)
