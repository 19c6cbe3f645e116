"""Transform chain that feeds the model — mirror of the reference's
core/datasets/torch_transforms.py (ToTensor :9-13, ToFullDense :17-40, Voxelization :44-81).
"""
from typing import Tuple

import numpy as np
import torch

from ...utils import voxelization as Vox


class ToTensor:
    """numpy -> float64 torch tensors (the reference's `astype(np.float)` is float64)."""

    def __call__(self, sample):
        return tuple(s if torch.is_tensor(s) else torch.from_numpy(np.asarray(s).astype(np.float64)) for s in sample)


class ToFullDense:
    """Regression grids -> belief grids: any voxel > 0 becomes 1 (same dtype)."""

    def __init__(self, apply=[True, True]) -> None:
        self.apply = apply

    def densify(self, tensor: torch.Tensor):
        return (tensor > 0).to(tensor)

    def __call__(self, sample):
        vox, gt = [self.densify(t) if self.apply[i] else t for i, t in enumerate(sample)]
        return vox, gt


class Voxelization:
    """Voxelizes raw (N,3) points + (N,) labels into (density[1,Z,X,Y], keep-fraction[1,Z,X,Y]).

    `vox_size` (voxel edge lengths) takes priority over `vxg_size` (grid size), as in the reference.
    `device_out=True` (extension) keeps the grids on the GPU as torch tensors.
    """

    def __init__(self, keep_labels, vox_size: Tuple[int] = None, vxg_size: Tuple[int] = None, device_out=False) -> None:
        if vox_size is None and vxg_size is None:
            raise ValueError("Voxel size or Voxelgrid size must be provided")
        self.vox_size = vox_size
        self.vxg_size = vxg_size
        self.keep_labels = keep_labels
        self.device_out = device_out

    def __call__(self, sample):
        pts, labels = sample
        d, f = Vox.voxelize_sample(pts, labels, self.keep_labels, voxelgrid_dims=self.vxg_size, voxel_dims=self.vox_size,
                                   device_out=self.device_out)
        return d[None], f[None]
