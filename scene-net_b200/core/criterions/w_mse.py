"""Weighted MSE criterion — drop-in mirror of the reference's core/criterions/w_mse.py (WeightedMSE :22-151):
same constructor, attributes (`freqs`, `ranges`, `weight_alpha`, `weight_epsilon`, `mse_weight`), methods and
`forward(y_pred, y_gt)`, but the ~25 full-tensor passes of `get_weight_target` + the weighted mean are ONE
reduction kernel (csrc/criterion.cu) and the backward is one elementwise kernel.

The weighting scheme is a 10-entry table: everything `get_dens_target` / `get_weight_target` do to a voxel
depends only on its histogram bin, so the reference's own tensor ops are run once on the bin indices (host,
10 elements, at construction) and the kernels look the result up per voxel.
"""
from __future__ import annotations

import os

import cloudpickle
import torch

from ... import ops

HIST_PATH = os.path.join(os.getcwd(), 'hist_estimation.pickle')


def save_pickle(data, filename):
    with open(filename, 'wb') as handle:
        cloudpickle.dump(data, handle)


def load_pickle(filename):
    """the shipped hist_estimation.pickle holds CUDA tensors: on a host without a device they are
    mapped to the CPU instead of failing"""
    with open(filename, 'rb') as handle:
        if torch.cuda.is_available():
            return cloudpickle.load(handle)
        orig = torch.storage._load_from_bytes
        torch.storage._load_from_bytes = lambda b: torch.load(__import__("io").BytesIO(b), map_location="cpu", weights_only=False)
        try:
            return cloudpickle.load(handle)
        finally:
            torch.storage._load_from_bytes = orig


class _FusedCriterion(torch.autograd.Function):
    """loss = [mse_weight * mean(w(y) (y - p)^2)] + [(1 - Tversky)^gamma]  (terms selected by spec.terms)"""

    @staticmethod
    def forward(ctx, pred, y, spec):
        loss, coef, p, t = ops.criterion_fwd(pred.detach(), y.detach(), spec)
        ctx.spec = spec
        ctx.shape = pred.shape
        ctx.save_for_backward(p, t, coef)
        return loss if pred.dtype == torch.float64 else loss.to(pred.dtype)

    @staticmethod
    def backward(ctx, g):
        p, t, coef = ctx.saved_tensors
        d = ops.criterion_bwd(p, t, coef, ctx.spec, grad_out=g)
        if d.shape != ctx.shape:  # pred was broadcast against y
            d = d.sum_to_size(ctx.shape)
        return d, None, None


class WeightedMSE(torch.nn.Module):

    def __init__(self, targets=None, weighting_scheme_path=HIST_PATH, weight_alpha=1, weight_epsilon=0.1, mse_weight=1, hist=None, **kwargs) -> None:
        """Same arguments as the reference (w_mse.py:24).  Extension: `hist=(freqs, ranges)` hands the
        histogram over directly (two 1-D tensors / sequences) instead of a pickle path."""
        super(WeightedMSE, self).__init__()
        self.weight_alpha = weight_alpha
        self.weight_epsilon = weight_epsilon
        self.mse_weight = mse_weight
        self.relu = torch.nn.ReLU()
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

        if hist is not None:
            self.freqs, self.ranges = torch.as_tensor(hist[0], dtype=torch.int64), torch.as_tensor(hist[1], dtype=torch.float32)
        elif weighting_scheme_path is not None and os.path.exists(weighting_scheme_path):
            self.pik_name = weighting_scheme_path
            self.freqs, self.ranges = load_pickle(self.pik_name)
        elif targets is not None:
            print("calculating histogram estimation...")
            self.freqs, self.ranges = self.hist_frequency_estimation(torch.flatten(targets), plot=False)
            save_pickle((self.freqs, self.ranges), f"{os.path.join('.', 'hist_estimation.pickle')}")
        else:
            raise ValueError("No targets were provided to build the weighting scheme")
        self.freqs = self.freqs.to(self.device)
        self.ranges = self.ranges.to(self.device)
        self._table_key = None
        self._table = None

    def hist_frequency_estimation(self, y: torch.Tensor, hist_len=10, plot=False):
        """(counts, bin starts) of y over `hist_len` equal bins of [0, 1) — setup time, not on the hot path
        (w_mse.py:71-112)."""
        y = y.to(self.device)
        starts = torch.linspace(0, 1, hist_len + 1, device=self.device)[:hist_len]
        counts = torch.bincount((hist_len * y).to(torch.int), minlength=hist_len)
        if plot:
            width = float(starts[1] - starts[0]) if hist_len > 1 else 1.0
            print("Histogram Bin /\t Count")
            for lo, c in zip(starts.tolist(), counts.tolist()):
                print(f"[{lo:.3f}, {lo + width:.3f}[ : {c}")
        return counts, starts

    # ------------------------------------------------------------------ the 10-entry weighting table
    def _weight_table(self):
        """(ranges, w_raw) as python floats: the reference's get_dens_target + the max() of get_weight_target
        (w_mse.py:114-143) applied to the bin indices themselves — same ops, same dtypes, 10 elements."""
        key = (self.freqs.data_ptr(), self.freqs._version, self.ranges.data_ptr(), self.ranges._version,
               float(self.weight_alpha), float(self.weight_epsilon))
        if self._table_key != key:
            freqs, ranges = self.freqs.detach().cpu(), self.ranges.detach().cpu()
            hist_idx = torch.arange(len(ranges), dtype=torch.int64)
            for idx in range(len(freqs)):
                hist_idx[hist_idx == idx] = freqs[idx]          # the reference's in-place replacement loop
            freq_min, freq_max = torch.min(freqs), torch.max(freqs)
            dens = (hist_idx - freq_min) / (freq_max - freq_min)  # int64 / int64 -> float32
            w = torch.max(1 - self.weight_alpha * dens, torch.full_like(dens, self.weight_epsilon))
            self._table = ([float(v) for v in ranges.to(torch.float32)], [float(v) for v in w.to(torch.float32)])
            self._dens_table, self._w_table = dens, w.to(torch.float32)
            self._table_key = key
        return self._table

    def _spec(self, terms: int, **tversky) -> ops.CriterionSpec:
        ranges, w_raw = self._weight_table()
        return ops.CriterionSpec(ranges=ranges, w_raw=w_raw, mse_weight=float(self.mse_weight), terms=terms, **tversky)

    def _bins(self, y: torch.Tensor) -> torch.Tensor:
        """nearest histogram bin of every target value (first minimum, like torch.argmin in w_mse.py:120)"""
        return (y.unsqueeze(-1) - self.ranges.to(y.device)).abs().argmin(dim=-1)

    def get_dens_target(self, y: torch.Tensor, calc_weights=False):
        """per-voxel density of the target's histogram bin (w_mse.py:114-133; diagnostic API, the fused forward never
        materialises it): a lookup of the bin in the 10-entry table the reference's own ops produce"""
        if calc_weights:
            self.freqs, self.ranges = self.hist_frequency_estimation(y)
            self._table_key = None
        self._weight_table()
        return self._dens_table.to(y.device)[self._bins(y)]

    def get_weight_target(self, y: torch.Tensor):
        """per-voxel weight max(1 - alpha * density, epsilon) / mean (w_mse.py:135-145; diagnostic API)"""
        y = y.to(self.device)
        self._weight_table()
        w = self._w_table.to(y.device)[self._bins(y)]
        return w / w.mean()

    def forward(self, y_pred: torch.Tensor, y_gt: torch.Tensor):
        return _FusedCriterion.apply(y_pred, y_gt, self._spec(1))

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = parent_parser.add_argument_group('WeightedMSE')
        parser.add_argument('--weight_alpha', type=float, default=1)
        parser.add_argument('--weight_epsilon', type=float, default=0.01)
        parser.add_argument('--mse_weight', type=float, default=1)
        parser.add_argument('--hist_path', type=str, default=HIST_PATH)
        return parent_parser
