"""Tversky / focal Tversky losses behind the reference's class names (core/criterions/tversky_loss.py: TverskyLoss
:10-53, FocalTverskyLoss :66-95).

    Tv = (TP + s) / (TP + alpha FP + beta FN + s),   TP = sum p y,  FP = sum (1 - y) p,  FN = sum y (1 - p)
    TverskyLoss = 1 - Tv,   FocalTverskyLoss = (1 - Tv) ** gamma

TP / FP / FN over the whole batch come out of ONE pass of the fused reduction kernel (csrc/criterion.cu) instead of
six full-tensor passes; the backward is one elementwise kernel with the closed-form dL/dp.  Both classes are thin
parameter holders around the same autograd function (gamma = 1 gives the plain Tversky loss).
"""
from __future__ import annotations

import torch.nn as nn

from ... import ops
from .w_mse import _FusedCriterion

ALPHA, BETA, GAMMA = 0.5, 1, 2

# (flag, default, help) of the command-line arguments the reference registers for these losses
_ARGS = {
    "tversky_alpha": (ALPHA, "controls the penalty for false positives"),
    "tversky_beta": (BETA, "controls the penalty for false negatives"),
    "tversky_smooth": (1, "smooth factor to avoid division by zero"),
    "focal_gamma": (GAMMA, "controls the penalty for easy examples"),
}


def _register(parent_parser, group_name, names):
    group = parent_parser.add_argument_group(group_name)
    for name in names:
        default, text = _ARGS[name]
        group.add_argument(f"--{name}", type=float, default=default, help=text)
    return parent_parser


class _TverskyFamily(nn.Module):
    """holds (alpha, beta, smooth, gamma) under the attribute names of the reference and evaluates the fused kernel"""

    def __init__(self, alpha, beta, smooth, gamma):
        super().__init__()
        self.tversky_alpha, self.tversky_beta, self.tversky_smooth = alpha, beta, smooth
        self._gamma = gamma

    def _exponent(self):
        return self._gamma

    def forward(self, inputs, targets):
        # a one-bin weighting table: the weighted-MSE term of the fused kernel is switched off (terms = 2)
        spec = ops.CriterionSpec(ranges=[0.0], w_raw=[1.0], terms=2, tversky_alpha=float(self.tversky_alpha),
                                 tversky_beta=float(self.tversky_beta), focal_gamma=float(self._exponent()),
                                 tversky_smooth=float(self.tversky_smooth))
        return _FusedCriterion.apply(inputs, targets, spec)


class TverskyLoss(_TverskyFamily):

    def __init__(self, tversky_alpha=ALPHA, tversky_beta=BETA, tversky_smooth=1, **kwargs):
        super().__init__(tversky_alpha, tversky_beta, tversky_smooth, 1.0)

    @staticmethod
    def add_model_specific_args(parent_parser):
        return _register(parent_parser, "TverskyLoss", ("tversky_alpha", "tversky_beta", "tversky_smooth"))


class FocalTverskyLoss(_TverskyFamily):

    def __init__(self, tversky_alpha=ALPHA, tversky_beta=BETA, focal_gamma=GAMMA, tversky_smooth=1, **kwargs):
        super().__init__(tversky_alpha, tversky_beta, tversky_smooth, focal_gamma)
        self.focal_gamma = focal_gamma

    def _exponent(self):
        return self.focal_gamma  # read at call time: the attribute may be reassigned like in the reference

    @staticmethod
    def add_model_specific_args(parent_parser):
        return _register(parent_parser, "FocalTverskyLoss", ("tversky_alpha", "tversky_beta", "tversky_smooth", "focal_gamma"))
