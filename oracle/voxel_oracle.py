"""TEST INFRASTRUCTURE ONLY — CPU (numpy) restatement of SCENE-Net's voxelization path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file; the product path (scene-net_b200/) never does.

What it restates (reference file:line, relative to the reference root):
  * bin edges + voxel indices ......... utils/pcd_processing.py:341-372 -> pyntcloud==0.1.6
                                        `structures/voxelgrid.py::VoxelGrid.compute`
  * count grid + per-column MinMax .... utils/voxelization.py:164-204, utils/pcd_processing.py:305-321
  * per-voxel max label ............... utils/voxelization.py:207-241
  * keep-label fraction ............... utils/voxelization.py:244-300
  * threshold ......................... utils/voxelization.py:304-323
  * transform glue / densify .......... core/datasets/torch_transforms.py:9-40,74-81

PARITY UNPINNED for the pyntcloud step: pyntcloud is a third-party dependency that is
neither vendored in the reference nor installed in the build image (requirements.txt:9,
`pyntcloud==0.1.6`).  `voxelgrid_compute` restates the published algorithm of
`VoxelGrid.compute` (bounding box -> cube -> np.linspace edges -> np.searchsorted - 1 ->
clip); no test or fixture in the reference pins results at that boundary.  Everything
downstream of the voxel indices (counts, max label, fractions, normalisation) IS pinned:
`tests/test_oracle_voxel.py` runs the reference's own `reg_on_voxel` (the one function of
the three that still executes under numpy 2 / pandas 3) through `PyntCloudShim`, checks
`normalize` against sklearn's MinMaxScaler and compares with the committed fixtures.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# pyntcloud 0.1.6 VoxelGrid.compute, restated
# --------------------------------------------------------------------------------------
def voxelgrid_compute(points, n_x=1, n_y=1, n_z=1, size_x=None, size_y=None, size_z=None,
                      regular_bounding_box=True):
    """Returns dict(xyzmin, xyzmax, x_y_z, segments, voxel_x, voxel_y, voxel_z).

    points: [N,3] float64 (x, y, z).  Index rule: voxel = clip(searchsorted(edges, p,
    side='left') - 1, 0, n): a point on an interior edge goes to the LOWER voxel, the
    minimum point to 0, the maximum point to n-1.
    """
    pts = np.asarray(points)
    x_y_z = [n_x, n_y, n_z]
    sizes = [size_x, size_y, size_z]
    xyzmin = pts.min(0)
    xyzmax = pts.max(0)
    xyz_range = np.ptp(pts, 0)
    if regular_bounding_box:
        margin = max(xyz_range) - xyz_range
        xyzmin = xyzmin - margin / 2
        xyzmax = xyzmax + margin / 2
    for n, size in enumerate(sizes):
        if size is None:
            continue
        margin = (((xyz_range[n] // size) + 1) * size) - xyz_range[n]
        xyzmin[n] -= margin / 2
        xyzmax[n] += margin / 2
        x_y_z[n] = int(((xyzmax[n] - xyzmin[n]) / size))
    segments = []
    shape = []
    for i in range(3):
        s, step = np.linspace(xyzmin[i], xyzmax[i], num=(x_y_z[i] + 1), retstep=True)
        segments.append(s)
        shape.append(step)
    vx = np.clip(np.searchsorted(segments[0], pts[:, 0]) - 1, 0, x_y_z[0])
    vy = np.clip(np.searchsorted(segments[1], pts[:, 1]) - 1, 0, x_y_z[1])
    vz = np.clip(np.searchsorted(segments[2], pts[:, 2]) - 1, 0, x_y_z[2])
    return dict(xyzmin=xyzmin, xyzmax=xyzmax, x_y_z=x_y_z, segments=segments, shape=shape,
                voxel_x=vx, voxel_y=vy, voxel_z=vz)


def linspace_edges(lo: float, hi: float, n: int) -> np.ndarray:
    """np.linspace(lo, hi, n+1) restated op by op: step = (hi-lo)/n (one rounding),
    e[j] = fl(fl(j*step) + lo) (two roundings, NO fma), e[n] = hi exactly.
    (numpy/_core/function_base.py::linspace; the `step == 0` branch divides first.)"""
    lo = np.float64(lo)
    hi = np.float64(hi)
    delta = hi - lo
    j = np.arange(0, n + 1, dtype=np.float64)
    step = delta / np.float64(n)
    if step == 0:
        e = (j / np.float64(n)) * delta
    else:
        e = j * step
    e = e + lo
    if n + 1 > 1:
        e[-1] = hi
    return e


def lin_index(vg, grid_zxy=None):
    """Linear voxel id in the reference grid layout data[z, x, y] (voxelization.py:193-200)."""
    nx, ny, nz = vg["x_y_z"]
    return (vg["voxel_z"].astype(np.int64) * nx + vg["voxel_x"]) * ny + vg["voxel_y"]


# --------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------
def _vg(xyz, voxelgrid_dims, voxel_dims):
    # utils/pcd_processing.py:360-368
    if voxel_dims is None:
        x, y, z = voxelgrid_dims
        return voxelgrid_compute(xyz, n_x=x, n_y=y, n_z=z)
    x, y, z = voxel_dims
    return voxelgrid_compute(xyz, size_x=x, size_y=y, size_z=z)


def raw_grids(xyz, labels=None, keep_labels=None, voxelgrid_dims=(64, 64, 64), voxel_dims=None):
    """count / keep-count / max-label grids, shape [nz, nx, ny] (int64, int64, float64)."""
    vg = _vg(xyz, voxelgrid_dims, voxel_dims)
    nx, ny, nz = vg["x_y_z"]
    lin = lin_index(vg)
    V = nz * nx * ny
    count = np.bincount(lin, minlength=V).reshape(nz, nx, ny)
    out = dict(vg=vg, lin=lin, count=count)
    if labels is not None:
        labels = np.asarray(labels, dtype=np.float64)
        if keep_labels is not None:
            keep = np.isin(labels, np.array(keep_labels).reshape(-1))
            out["keep"] = np.bincount(lin, weights=keep.astype(np.float64), minlength=V) \
                .astype(np.int64).reshape(nz, nx, ny)
        mx = np.full(V, -np.inf)
        np.maximum.at(mx, lin, labels)
        mx[count.reshape(-1) == 0] = 0.0
        out["maxlab"] = mx.reshape(nz, nx, ny)
    return out


def normalize_minmax(data: np.ndarray) -> np.ndarray:
    """sklearn MinMaxScaler().fit_transform(data.reshape(-1, ny)) restated
    (utils/pcd_processing.py:305-321): per-y-column min/max over all (z, x) rows;
    scale = 1/(max-min) (1 where the range is 0); out = data*scale + (0 - min*scale)."""
    shp = data.shape
    d = data.reshape(-1, shp[-1]).astype(np.float64)
    dmin = d.min(0)
    dmax = d.max(0)
    rng = dmax - dmin
    rng[rng == 0.0] = 1.0
    scale = 1.0 / rng
    mn = 0.0 - dmin * scale
    out = d * scale
    out += mn
    return out.reshape(shp)


def hist_on_voxel(xyz, voxelgrid_dims=(64, 64, 64), voxel_dims=None):
    g = raw_grids(xyz, voxelgrid_dims=voxelgrid_dims, voxel_dims=voxel_dims)
    return normalize_minmax(g["count"].astype(np.float64))


def classes_on_voxel(xyz, labels, voxel_dims=(64, 64, 64)):
    # NB the reference passes its `voxel_dims` argument as voxelgrid_dims (voxelization.py:229)
    g = raw_grids(xyz, labels, voxelgrid_dims=voxel_dims)
    return g["maxlab"]


def reg_on_voxel(xyz, labels, tower_label, voxelgrid_dims=(64, 64, 64), voxel_dims=None):
    g = raw_grids(xyz, labels, tower_label, voxelgrid_dims, voxel_dims)
    count = g["count"].astype(np.float64)
    keep = g["keep"].astype(np.float64)
    out = np.zeros_like(count)
    m = count > 0
    out[m] = keep[m] / count[m]
    return out


def prob_to_label(grid, tau):
    return (grid >= tau).astype(grid.dtype)


def voxelization_transform(pts, labels, keep_labels, vox_size=None, vxg_size=None):
    """Voxelization.__call__ (torch_transforms.py:74-81): (density[None], frac[None])."""
    d = hist_on_voxel(pts, voxel_dims=vox_size, voxelgrid_dims=vxg_size)
    f = reg_on_voxel(pts, labels, keep_labels, voxel_dims=vox_size, voxelgrid_dims=vxg_size)
    return d[None], f[None]


def densify(a):
    """ToFullDense.densify (torch_transforms.py:33-34)."""
    return (a > 0).astype(a.dtype)


# --------------------------------------------------------------------------------------
# functional stand-ins so the REAL reference functions can run in the build container
# --------------------------------------------------------------------------------------
class O3DPointCloudShim:
    def __init__(self):
        self.points = None


class _VoxelGridShim:
    def __init__(self, vg):
        self.x_y_z = vg["x_y_z"]
        self.voxel_x = vg["voxel_x"]
        self.voxel_y = vg["voxel_y"]
        self.voxel_z = vg["voxel_z"]
        self.segments = vg["segments"]
        self.shape = vg["shape"]
        self.xyzmin = vg["xyzmin"]
        self.xyzmax = vg["xyzmax"]


class PyntCloudShim:
    """Implements only what utils/pcd_processing.py:360-368 touches."""

    def __init__(self, points):
        self.xyz = np.asarray(points, dtype=np.float64)
        self.structures = {}

    @classmethod
    def from_instance(cls, library, instance):
        assert library == "open3d"
        return cls(np.asarray(instance.points))

    def add_structure(self, name, **kwargs):
        assert name == "voxelgrid"
        vg = voxelgrid_compute(self.xyz, **kwargs)
        vid = "V({},{},{})".format(vg["x_y_z"], [kwargs.get("size_x"), kwargs.get("size_y"), kwargs.get("size_z")], True)
        self.structures[vid] = _VoxelGridShim(vg)
        return vid


def vxg_to_xyz(vxg, origin=None, voxel_size=None) -> np.ndarray:
    """utils/voxelization.py:328-360: every voxel of a 3-D grid as a row (origin + index * voxel_size, value), C order
    of the index (np.indices(shape).reshape(3, -1).T), float64."""
    a = np.asarray(vxg)
    origin = np.array([0, 0, 0]) if origin is None else np.asarray(origin)
    voxel_size = np.array([1, 1, 1]) if voxel_size is None else np.asarray(voxel_size)
    idx = np.indices(a.shape).reshape(3, -1).T
    pts = origin + idx * voxel_size
    return np.concatenate((pts, a.reshape(-1, 1)), axis=1).astype(np.float64)
