"""Voxelization functions — drop-in mirror of the hot functions of the reference's
utils/voxelization.py (hist_on_voxel :164-204, classes_on_voxel :207-241, reg_on_voxel
:244-300, prob_to_label :304-323): numpy in, numpy out, same signatures and grid layout
[z, x, y] — computed by the CUDA kernels of csrc/voxelize.cu (bounding box, linspace edges,
searchsorted binning with warp-aggregated atomics, finalize) instead of pyntcloud + pandas
groupby + Python loops.  Plotting helpers of the reference are out of scope.

`voxelize_sample` does what the reference's Voxelization transform needs (density AND keep
fraction) in ONE pass over the points; the reference bins every cloud twice.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from .. import ops, voxel_ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("scenenet_b200 voxelization runs on the GPU; no CUDA device is available and there is no "
                           "CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(a, dev) -> torch.Tensor:
    if torch.is_tensor(a):
        return a.to(device=dev, dtype=torch.float64)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def _grids(xyz, labels, keep_labels, voxelgrid_dims, voxel_dims, want):
    dev = _device()
    pts = _to_device(xyz, dev)
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError(f"xyz must be (N, 3), got {tuple(pts.shape)}")
    if pts.stride(1) != 1:
        pts = pts.contiguous()
    lab = None if labels is None else _to_device(labels, dev).reshape(-1)
    edges = None
    if voxel_dims is not None:  # size mode overrides voxelgrid_dims (pcd_processing.py:364-367)
        off = torch.tensor([0, pts.shape[0]], dtype=torch.int64, device=dev)
        mnmx = voxel_ops.bounding_boxes(pts, off).cpu().numpy()[0]
        grid, e = voxel_ops.size_mode_edges(mnmx, voxel_dims)
        edges = torch.from_numpy(e).to(dev).reshape(1, -1)
    else:
        grid = tuple(int(v) for v in voxelgrid_dims)
    return voxel_ops.voxelize_clouds(pts, None, grid, lab, keep_labels, want=want, occ_dtype=torch.float64, edges=edges)


def hist_on_voxel(xyz, voxelgrid_dims=(64, 64, 64), voxel_dims=None):
    """Point count per voxel, MinMax-normalised per y column; ndarray [n_z, n_x, n_y] float64."""
    return _grids(xyz, None, None, voxelgrid_dims, voxel_dims, ("density",))["density"][0].cpu().numpy()


def classes_on_voxel(xyz, labels, voxel_dims=(64, 64, 64)):
    """Per-voxel maximum label (0 where empty).  NB: as in the reference (:229) the argument named
    `voxel_dims` is the voxel GRID size."""
    return _grids(xyz, labels, None, voxel_dims, None, ("max_label",))["max_label"][0].cpu().numpy()


def reg_on_voxel(xyz, labels, tower_label, voxelgrid_dims=(64, 64, 64), voxel_dims=None):
    """Per-voxel fraction of points whose label is in `tower_label` (int or list)."""
    return _grids(xyz, labels, tower_label, voxelgrid_dims, voxel_dims, ("frac",))["frac"][0].cpu().numpy()


def voxelize_sample(xyz, labels, keep_labels, voxelgrid_dims=(64, 64, 64), voxel_dims=None, device_out=False):
    """(density, frac) of one cloud in a single pass."""
    g = _grids(xyz, labels, keep_labels, voxelgrid_dims, voxel_dims, ("density", "frac"))
    if device_out:
        return g["density"][0], g["frac"][0]
    return g["density"][0].cpu().numpy(), g["frac"][0].cpu().numpy()


def prob_to_label(voxelgrid: Union[torch.Tensor, np.ndarray], tau: float) -> Union[torch.Tensor, np.ndarray]:
    """(voxelgrid >= tau) as 0/1 in the input's dtype."""
    if isinstance(voxelgrid, torch.Tensor):
        if voxelgrid.is_cuda and voxelgrid.dtype in (torch.float32, torch.float64):
            return ops.threshold(voxelgrid, tau)
        dev = _device()
        return ops.threshold(voxelgrid.to(dev, torch.float64), tau).to(device=voxelgrid.device, dtype=voxelgrid.dtype)
    a = np.asarray(voxelgrid)
    t = ops.threshold(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(_device()), tau)
    return t.cpu().numpy().astype(a.dtype)


def vxg_to_xyz(vxg: Union[torch.Tensor, np.ndarray], origin=None, voxel_size=None) -> np.ndarray:
    """utils/voxelization.py:328-360 of the reference: a voxel grid as a raw point cloud, (N, 4) numpy float64 with
    N = vxg.numel(): rows (origin + index * voxel_size, vxg[index]) for EVERY voxel in C order (callers keep the rows with
    label 1).  The reference loops over the voxels in Python (one tensor index per voxel: ~3 s for a 64^3 grid); here
    one kernel writes the rows.  Defaults as there: origin (0, 0, 0), voxel_size (1, 1, 1)."""
    origin = (0.0, 0.0, 0.0) if origin is None else np.asarray(origin, dtype=np.float64).reshape(3)
    voxel_size = (1.0, 1.0, 1.0) if voxel_size is None else np.asarray(voxel_size, dtype=np.float64).reshape(3)
    t = vxg if isinstance(vxg, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vxg))
    if not t.is_cuda:
        t = t.to(_device())
    return ops.vxg_to_xyz(t, origin, voxel_size).cpu().numpy()
