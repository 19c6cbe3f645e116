"""TEST INFRASTRUCTURE ONLY — CPU restatement of the metric collection the reference updates every step
(utils/scripts_utils.py:80-91, fed in core/lit_modules/lit_model_wrappers.py:170-171 / 189-190 / 199-200):

    MetricCollection([JaccardIndex(num_classes=2, threshold=tau), Precision(threshold=tau), Recall(threshold=tau),
                      F1Score(threshold=tau), FBetaScore(beta=0.5, threshold=tau)])

called as `metrics(torch.flatten(pred), torch.flatten(y).to(torch.int))`, `metrics.compute()`, `metrics.reset()`.

The arithmetic lives in a THIRD-PARTY dependency that is absent from /root/reference and from this image:
**torchmetrics == 0.9.0** (requirements.txt:17; the recorded run's files/requirements.txt:288).  This file restates the
published algorithm of that version for the binary input case the reference produces (float predictions in [0, 1),
integer targets in {0, 1}; `_input_format_classification` -> `preds >= threshold`):

* Precision / Recall / F1Score / FBetaScore are `StatScores` with `average='micro'`, `mdmc_average=None`: state =
  (tp, fp, tn, fn) summed over updates; `_reduce_stat_scores`: float32 `numerator / denominator`, 0 where the
  denominator is 0 (`zero_division = 0`).  precision = tp / (tp + fp), recall = tp / (tp + fn);
  `_fbeta_compute`: (1 + b^2) * precision * recall / (b^2 * precision + recall) (denominator 0 -> 1), all float32.
* JaccardIndex(num_classes=2) is a `ConfusionMatrix`: confmat[target, pred]; `_jaccard_from_confmat`:
  intersection = diag, union = row sums + column sums - diag, score_c = intersection_c / union_c
  (`absent_score = 0.0` where union_c == 0), result = mean over the two classes ('elementwise_mean').
* `MetricCollection.forward` returns the values of THIS batch under the metric class names and accumulates the state;
  `compute()` evaluates the accumulated state.

PARITY PARTLY PINNED: torchmetrics itself cannot be run here (not installed, no network), so the restatement is held
to scikit-learn's `precision_score` / `recall_score` / `f1_score` / `fbeta_score` / `jaccard_score(average='macro')`
on thresholded inputs (tests/test_oracle_metrics.py) — the same published definitions, an independent
implementation; the float32 rounding order of torchmetrics 0.9.0 is restated from its source, unpinned.
Only `tests/` may import this file.
"""
from __future__ import annotations

import numpy as np

NAMES = ("JaccardIndex", "Precision", "Recall", "F1Score", "FBetaScore")


def confusion_counts(pred: np.ndarray, y: np.ndarray, tau: float) -> np.ndarray:
    """[tp, fp, tn, fn] of (pred >= tau) vs (y != 0) — `_input_format_classification` + `_stat_scores`."""
    pred = np.asarray(pred).reshape(-1)
    tgt = np.asarray(y).reshape(-1) != 0
    thr = np.float32(tau) if pred.dtype == np.float32 else np.float64(tau)
    pos = pred >= thr
    return np.array([np.sum(pos & tgt), np.sum(pos & ~tgt), np.sum(~pos & ~tgt), np.sum(~pos & tgt)], dtype=np.int64)


def _safe_div32(num, den) -> np.float32:
    num, den = np.float32(num), np.float32(den)
    return np.float32(0.0) if den == 0 else np.float32(num / den)


def values_from_counts(counts) -> dict:
    """the five metric values (float32) from accumulated [tp, fp, tn, fn]"""
    tp, fp, tn, fn = (int(v) for v in counts)
    precision = _safe_div32(tp, tp + fp)
    recall = _safe_div32(tp, tp + fn)

    def fbeta(beta: float) -> np.float32:
        b2 = beta ** 2
        num = np.float32(np.float32(1 + b2) * precision) * recall
        den = np.float32(np.float32(b2) * precision) + recall
        if den == 0:
            den = np.float32(1.0)
        return np.float32(num / den)

    # confusion matrix rows = target, columns = prediction: [[tn, fp], [fn, tp]]
    inter = (tn, tp)
    union = (tn + fp + fn, tp + fp + fn)
    iou = [np.float32(0.0) if u == 0 else np.float32(np.float32(i) / np.float32(u)) for i, u in zip(inter, union)]
    jacc = np.float32((iou[0] + iou[1]) / np.float32(2))
    return {"JaccardIndex": jacc, "Precision": precision, "Recall": recall, "F1Score": fbeta(1.0), "FBetaScore": fbeta(0.5)}


class MetricCollectionOracle:
    """update / forward / compute / reset of the collection over a stream of (pred, y) batches"""

    def __init__(self, tau: float = 0.65):
        self.tau = tau
        self.state = np.zeros(4, dtype=np.int64)

    def __call__(self, pred, y) -> dict:
        c = confusion_counts(pred, y, self.tau)
        self.state += c
        return values_from_counts(c)

    def compute(self) -> dict:
        return values_from_counts(self.state)

    def reset(self) -> None:
        self.state[:] = 0
