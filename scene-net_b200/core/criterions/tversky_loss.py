"""Tversky / focal Tversky losses — drop-in mirror of the reference's core/criterions/tversky_loss.py
(TverskyLoss :10-53, FocalTverskyLoss :66-95): TP / FP / FN over the whole batch come from the fused reduction
kernel (csrc/criterion.cu) instead of six full-tensor passes, the backward is one elementwise kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from .w_mse import _FusedCriterion

ALPHA = 0.5
BETA = 1
GAMMA = 2


def _tversky_spec(alpha, beta, gamma, smooth) -> ops.CriterionSpec:
    # a one-bin weighting table: the weighted-MSE term is switched off (terms = 2)
    return ops.CriterionSpec(ranges=[0.0], w_raw=[1.0], terms=2, tversky_alpha=float(alpha), tversky_beta=float(beta),
                             focal_gamma=float(gamma), tversky_smooth=float(smooth))


class TverskyLoss(nn.Module):

    def __init__(self, tversky_alpha=ALPHA, tversky_beta=BETA, tversky_smooth=1, **kwargs):
        super(TverskyLoss, self).__init__()
        self.tversky_alpha = tversky_alpha
        self.tversky_beta = tversky_beta
        self.tversky_smooth = tversky_smooth

    def forward(self, inputs, targets):
        return _FusedCriterion.apply(inputs, targets, _tversky_spec(self.tversky_alpha, self.tversky_beta, 1.0, self.tversky_smooth))

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = parent_parser.add_argument_group('TverskyLoss')
        parser.add_argument('--tversky_alpha', type=float, default=ALPHA)
        parser.add_argument('--tversky_beta', type=float, default=BETA)
        parser.add_argument('--tversky_smooth', type=float, default=1)
        return parent_parser


class FocalTverskyLoss(nn.Module):

    def __init__(self, tversky_alpha=ALPHA, tversky_beta=BETA, focal_gamma=GAMMA, tversky_smooth=1, **kwargs):
        super(FocalTverskyLoss, self).__init__()
        self.tversky_alpha = tversky_alpha
        self.tversky_beta = tversky_beta
        self.tversky_smooth = tversky_smooth
        self.focal_gamma = focal_gamma

    def forward(self, inputs, targets):
        return _FusedCriterion.apply(inputs, targets, _tversky_spec(self.tversky_alpha, self.tversky_beta, self.focal_gamma,
                                                                    self.tversky_smooth))

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = parent_parser.add_argument_group('FocalTverskyLoss')
        parser.add_argument('--tversky_alpha', type=float, default=ALPHA, help='controls the penalty for false positives')
        parser.add_argument('--tversky_beta', type=float, default=BETA, help='controls the penalty for false negatives')
        parser.add_argument('--tversky_smooth', type=float, default=1, help='smooth factor to avoid division by zero')
        parser.add_argument('--focal_gamma', type=float, default=GAMMA, help='controls the penalty for easy examples')
        return parent_parser
