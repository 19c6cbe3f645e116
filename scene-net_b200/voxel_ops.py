"""GPU voxelization entry points (tensor level) over the C ABI.

`voxelize_clouds` is the B200-native data path: any number of clouds in ONE set of launches
(bounding boxes -> edges -> binning -> finalize), inputs and outputs resident in HBM.
The numpy-facing mirrors of the reference's functions live in utils/voxelization.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import SN_F32, SN_F64, check, lib
from .ops import _need_cuda, _on_device, _ptr, _stream

_keep_cache: dict = {}


def _keep_tensor(keep: Optional[Sequence[float]], device) -> Optional[torch.Tensor]:
    if keep is None:
        return None
    vals = tuple(float(v) for v in np.array(keep, dtype=np.float64).reshape(-1))
    key = (vals, device.index)
    t = _keep_cache.get(key)
    if t is None:
        t = torch.tensor(vals, dtype=torch.float64, device=device)
        _keep_cache[key] = t
    return t


def bounding_boxes(points: torch.Tensor, offsets: Optional[torch.Tensor]) -> torch.Tensor:
    """[C,6] float64 (xmin,ymin,zmin,xmax,ymax,zmax) per cloud (offsets None = one cloud)."""
    _need_cuda(points, "points")
    C_ = 1 if offsets is None else offsets.numel() - 1
    out = torch.empty((C_, 6), dtype=torch.float64, device=points.device)
    with _on_device(points.device):
        check(lib.sn_vox_minmax(points.data_ptr(), points.stride(0), _ptr(offsets), C_, points.shape[0], out.data_ptr(), _stream()),
              "sn_vox_minmax")
    return out


def grid_edges(mnmx: torch.Tensor, grid_xyz: Sequence[int]) -> torch.Tensor:
    nx, ny, nz = (int(v) for v in grid_xyz)
    C_ = mnmx.shape[0]
    out = torch.empty((C_, nx + ny + nz + 3), dtype=torch.float64, device=mnmx.device)
    with _on_device(mnmx.device):
        check(lib.sn_vox_edges(mnmx.data_ptr(), C_, nx, ny, nz, out.data_ptr(), _stream()), "sn_vox_edges")
    return out


def voxelize_clouds(points: torch.Tensor, offsets: Optional[torch.Tensor] = None, grid_xyz: Sequence[int] = (64, 64, 64),
                    labels: Optional[torch.Tensor] = None, keep_labels: Optional[Sequence[float]] = None,
                    want=("density", "frac"), occ_dtype: torch.dtype = torch.float32, edges: Optional[torch.Tensor] = None,
                    return_lin: bool = False) -> dict:
    """points: [N, ld] float64 CUDA rows (x,y,z first; ld = points.stride(0) may exceed 3, e.g. the
    TS40K rows x,y,z,label).  offsets: [C+1] int64 CUDA (None = one cloud).  grid_xyz = (n_x,n_y,n_z)
    as in the reference's `voxelgrid_dims`; grids come back as [C, n_z, n_x, n_y].
    want: any of density, frac, max_label, occ, occ_keep, count, keep_count.
    """
    _need_cuda(points, "points")
    if points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] < 3 or points.stride(1) != 1:
        raise TypeError("points must be float64 [N, >=3] with unit inner stride")
    dev = points.device
    N = points.shape[0]
    C_ = 1 if offsets is None else offsets.numel() - 1  # one cloud: no offsets tensor (no host->device copy per call)
    nx, ny, nz = (int(v) for v in grid_xyz)
    want = set(want)
    if labels is not None:
        _need_cuda(labels, "labels")
        if labels.dtype != torch.float64 or labels.dim() != 1 or labels.shape[0] != N:
            raise TypeError("labels must be float64 [N]")
    need_keep = bool(want & {"frac", "occ_keep", "keep_count"})
    if need_keep and (labels is None or keep_labels is None):
        raise ValueError("frac / occ_keep need labels and keep_labels")
    keep_t = _keep_tensor(keep_labels, dev) if need_keep else None
    with _on_device(dev):
        shape = (C_, nz, nx, ny)
        count = torch.empty(shape, dtype=torch.int32, device=dev)
        keep_count = torch.empty(shape, dtype=torch.int32, device=dev) if need_keep else None
        maxlab = torch.empty(shape, dtype=torch.float64, device=dev) if "max_label" in want else None
        lin = torch.empty(N, dtype=torch.int32, device=dev) if return_lin else None
        lab_ld = labels.stride(0) if labels is not None else 0
        n_keep = 0 if keep_t is None else keep_t.numel()
        if edges is None:
            # bounding boxes -> edges -> binning in three launches (one init kernel, edges derived inside the binning kernel)
            mnmx = torch.empty((C_, 6), dtype=torch.float64, device=dev)
            edges = torch.empty((C_, nx + ny + nz + 3), dtype=torch.float64, device=dev)
            check(lib.sn_vox_voxelize(points.data_ptr(), points.stride(0), _ptr(labels), lab_ld, _ptr(offsets), C_, N, nx, ny, nz,
                                      _ptr(keep_t), n_keep, mnmx.data_ptr(), edges.data_ptr(), count.data_ptr(), _ptr(keep_count),
                                      _ptr(maxlab), _ptr(lin), _stream()), "sn_vox_voxelize")
        else:
            check(lib.sn_vox_bin(points.data_ptr(), points.stride(0), _ptr(labels), lab_ld, _ptr(offsets), C_, N, edges.data_ptr(),
                                 nx, ny, nz, _ptr(keep_t), n_keep, count.data_ptr(), _ptr(keep_count), _ptr(maxlab), _ptr(lin),
                                 _stream()), "sn_vox_bin")
        density = torch.empty(shape, dtype=torch.float64, device=dev) if "density" in want else None
        frac = torch.empty(shape, dtype=torch.float64, device=dev) if "frac" in want else None
        occ = torch.empty(shape, dtype=occ_dtype, device=dev) if "occ" in want else None
        occ_keep = torch.empty(shape, dtype=occ_dtype, device=dev) if "occ_keep" in want else None
        ws = None
        if density is not None:
            ws = torch.empty(int(lib.sn_vox_finalize_workspace_bytes(C_, ny)), dtype=torch.uint8, device=dev)
        if any(t is not None for t in (density, frac, maxlab, occ, occ_keep)):
            check(lib.sn_vox_finalize(count.data_ptr(), _ptr(keep_count), C_, nx, ny, nz, _ptr(density), _ptr(frac),
                                      _ptr(maxlab), _ptr(occ), _ptr(occ_keep),
                                      SN_F64 if occ_dtype == torch.float64 else SN_F32, _ptr(ws), _stream()),
                  "sn_vox_finalize")
    out = {"count": count, "edges": edges}
    for k, v in (("keep_count", keep_count), ("max_label", maxlab), ("density", density), ("frac", frac), ("occ", occ),
                 ("occ_keep", occ_keep), ("lin", lin)):
        if v is not None:
            out[k] = v
    return out


def size_mode_edges(mnmx_host: np.ndarray, voxel_dims: Sequence[float]):
    """Host-side grid definition for the reference's size mode (`voxel_dims`, pcd_processing.py:364-367
    -> pyntcloud VoxelGrid.compute with size_x/y/z): the grid extent depends on the data, so the
    bounding box is read back once; the binning itself stays on the GPU.
    Returns ((n_x,n_y,n_z), edges[(nx+1)+(ny+1)+(nz+1)])."""
    mn = np.array(mnmx_host[:3], dtype=np.float64)
    mx = np.array(mnmx_host[3:], dtype=np.float64)
    rng = mx - mn
    margin = max(rng) - rng
    mn = mn - margin / 2
    mx = mx + margin / 2
    n = [0, 0, 0]
    for k, size in enumerate(voxel_dims):
        m = (((rng[k] // size) + 1) * size) - rng[k]
        mn[k] -= m / 2
        mx[k] += m / 2
        n[k] = int((mx[k] - mn[k]) / size)
    edges = np.concatenate([np.linspace(mn[k], mx[k], n[k] + 1) for k in range(3)])
    return tuple(n), edges
