"""Metric collection of the training loop — mirror of `init_metrics` in the reference's utils/scripts_utils.py:80-91
(same in utils/observer_utils.py:221-232), which builds a torchmetrics 0.9.0 `MetricCollection` of
JaccardIndex(num_classes=2) / Precision / Recall / F1Score / FBetaScore(beta=0.5), all thresholded at tau, and is used
by core/lit_modules/lit_model_wrappers.py as

    self.train_metrics(torch.flatten(pred), torch.flatten(y).to(torch.int)).update()     # :170-171, every step
    metric_res = metrics.compute(); ...; metrics.reset()                                   # :122-130, every epoch

All five metrics are functions of the same four confusion counts, so here ONE kernel pass over (pred, y) per step
(csrc/metrics.cu, sn_confusion_counts) updates a shared int64 [4] device state {TP, FP, TN, FN}; the five values are
a few scalar float32 operations on it, evaluated only when somebody reads them (the training loop discards the
per-step return value).  The reference makes ~5 x (threshold + cast + 4 masked sums / bincount) full-tensor passes.
SURVEY §8(f) rank 2.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from .. import ops

METRIC_NAMES = ("JaccardIndex", "Precision", "Recall", "F1Score", "FBetaScore")


def _values(counts: torch.Tensor) -> "OrderedDict[str, torch.Tensor]":
    """the five float32 scalars from int64 [tp, fp, tn, fn] (device tensors, no host synchronisation); float32 operation
    order of torchmetrics 0.9.0 (`_reduce_stat_scores`, `_fbeta_compute`, `_jaccard_from_confmat`)"""
    tp, fp, tn, fn = counts.unbind(0)
    zero, one = torch.zeros((), dtype=torch.float32, device=counts.device), torch.ones((), dtype=torch.float32, device=counts.device)

    def safe_div(num, den):
        den = den.to(torch.float32)
        return torch.where(den == 0, zero, num.to(torch.float32) / torch.where(den == 0, one, den))

    precision, recall = safe_div(tp, tp + fp), safe_div(tp, tp + fn)

    def fbeta(beta: float):
        b2 = beta ** 2
        num = (1 + b2) * precision * recall
        den = b2 * precision + recall
        return num / torch.where(den == 0, one, den)

    jaccard = (safe_div(tn, tn + fp + fn) + safe_div(tp, tp + fp + fn)) / 2
    return OrderedDict((("JaccardIndex", jaccard), ("Precision", precision), ("Recall", recall), ("F1Score", fbeta(1.0)),
                        ("FBetaScore", fbeta(0.5))))


class _LazyValues(dict):
    """what `collection(preds, target)` returns: the batch values under the metric names, computed from the batch counts
    on first access.  `.update()` without arguments (what the reference calls on it) stays the no-op it is on a dict."""

    def __init__(self, counts: torch.Tensor, prefix: str):
        super().__init__()
        self._counts, self._prefix, self._done = counts, prefix, False

    def _fill(self):
        if not self._done:
            self._done = True
            for k, v in _values(self._counts).items():
                dict.__setitem__(self, self._prefix + k, v)

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        self._fill()
        return dict.__len__(self)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)


class ConfusionMetricCollection(torch.nn.Module):
    """Drop-in for the `MetricCollection` of `init_metrics`: `__call__(preds, target)` (accumulates, returns the batch
    values), `update`, `compute`, `reset`, `clone(prefix=...)`; the state travels with `.to(device)`."""

    def __init__(self, tau: float = 0.65, prefix: Optional[str] = None):
        super().__init__()
        self.threshold = float(tau)
        self.prefix = prefix or ""
        self.register_buffer("confusion", torch.zeros(4, dtype=torch.int64), persistent=False)  # tp, fp, tn, fn

    def _move_state(self, device):
        if self.confusion.device != device:
            self.confusion = self.confusion.to(device)

    @torch.no_grad()
    def update(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        self._move_state(preds.device)
        return ops.confusion_counts(preds.detach().reshape(-1), target.reshape(-1), self.threshold, total=self.confusion)

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> dict:
        return _LazyValues(self.update(preds, target), self.prefix)

    def compute(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((self.prefix + k, v) for k, v in _values(self.confusion).items())

    def reset(self) -> None:
        self.confusion.zero_()

    def clone(self, prefix: Optional[str] = None) -> "ConfusionMetricCollection":
        c = ConfusionMetricCollection(self.threshold, self.prefix if prefix is None else prefix)
        c.confusion = self.confusion.clone()
        return c

    def keys(self):
        return [self.prefix + k for k in METRIC_NAMES]


def init_metrics(tau=0.65):
    return ConfusionMetricCollection(tau)
