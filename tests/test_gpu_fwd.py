"""GPU parity of the two observer-forward kernels (dense stencil, occupancy-driven) and of the device-side
selection between them, through the C ABI (sn_grid_prepare, sn_scenenet_fwd).

Checker: relu(tanh(conv3d(x, K, padding='same'))) evaluated by torch in float64 on the same inputs
(SCENE_Net.py:324-337 with the G kernels already combined into one).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_fwd(x, K):
    ks = K.shape
    pl = [(k - 1) // 2 for k in ks]
    pr = [k - 1 - l for k, l in zip(ks, pl)]
    xp = F.pad(x.to(torch.float64), (pl[2], pr[2], pl[1], pr[1], pl[0], pr[0]))
    s = F.conv3d(xp, K.to(torch.float64)[None, None])
    return torch.relu(torch.tanh(s)), s


def _inputs(B, grid, ks, density, seed, binary):
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand((B, 1, *grid), generator=g) < density
    val = torch.ones((B, 1, *grid)) if binary else torch.rand((B, 1, *grid), generator=g) + 0.25
    x = (occ * val).to(torch.float32)
    K = (torch.randn(ks, generator=g) * 0.2).to(torch.float32)
    return x.to(DEV), K.to(DEV)


CASES = [
    # B, grid (Z,X,Y), kernel, density, binary
    (2, (64, 64, 64), (9, 5, 5), 0.016, True),      # config-2 shape
    (3, (20, 17, 23), (9, 5, 5), 0.05, False),      # ragged, Y % 4 != 0 -> plain-load path
    (2, (24, 24, 24), (9, 7, 7), 0.02, True),       # two tap groups per lane
    (1, (16, 16, 16), (3, 3, 3), 0.3, False),
    (1, (12, 12, 12), (5, 5, 4), 0.1, False),       # ky = 4: no dense instantiation
    (2, (16, 16, 16), (4, 6, 5), 0.02, False),      # even extents (asymmetric 'same' padding)
    (1, (32, 32, 32), (9, 9, 9), 0.016, True),      # three tap groups per lane
    (1, (32, 32, 32), (11, 11, 11), 0.016, True),   # four tap groups per lane
    (2, (8, 8, 128), (9, 5, 5), 0.016, True),       # several y tiles
    (1, (40, 8, 32), (1, 1, 1), 0.5, False),        # single tap
    (2, (64, 64, 64), (9, 5, 5), 0.0, True),        # empty grids
    (1, (32, 64, 64), (9, 5, 5), 1.0, False),       # full grids: 7 list rounds per row
    (1, (16, 64, 64), (9, 5, 5), 0.4, False),       # 3 list rounds
]


@pytest.mark.parametrize("B,grid,ks,density,binary", CASES)
@pytest.mark.parametrize("out_dtype", [torch.float64, torch.float32])
def test_sparse_and_dense_forward_match_float64_reference(B, grid, ks, density, binary, out_dtype):
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_PATH_DENSE, SN_PATH_SPARSE
    x, K = _inputs(B, grid, ks, density, seed=hash((B, grid, ks)) % 1000, binary=binary)
    ref, s = _ref_fwd(x, K)
    # float32 dot products: error bounded by the magnitude of the summed terms
    _, sabs = _ref_fwd(x.abs(), K.abs())
    tol = 4e-6 * sabs + 1e-7
    ps = ops.scenenet_fwd(x, K, out_dtype, mode=SN_PATH_SPARSE)
    pd = ops.scenenet_fwd(x, K, out_dtype, mode=SN_PATH_DENSE)
    assert ps.dtype == out_dtype and ps.shape == x.shape
    for name, p in (("sparse", ps), ("dense", pd)):
        err = (p.to(torch.float64) - ref).abs()
        assert bool((err <= tol).all()), f"{name}: max err {float(err.max()):.3e} (tol {float(tol.max()):.3e})"
    assert torch.equal(ps, ops.scenenet_fwd(x, K, out_dtype, mode=SN_PATH_SPARSE)), "deterministic"


@pytest.mark.parametrize("density,expect_sparse", [(0.005, True), (0.016, True), (0.045, True), (0.3, False), (1.0, False)])
def test_device_side_selection(density, expect_sparse):
    """AUTO + the state buffer of sn_grid_prepare: the kernel is chosen per tile on the device (ABI v4; a tile goes to the
    dense stencil when its halo box is more than ~10 % occupied) — uniform grids get one kernel for all tiles."""
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_PATH_AUTO, SN_PATH_DENSE, SN_PATH_SPARSE
    ks = (9, 5, 5)
    x, K = _inputs(2, (32, 32, 64), ks, density, seed=7, binary=False)
    x32, nnz = ops.prepare(x.to(torch.float64))
    pa = ops.scenenet_fwd(x32, K, torch.float64, nnz=nnz, mode=SN_PATH_AUTO)
    # forced modes with the state buffer: the same kernels AUTO chooses between (mask-driven / dense)
    pe = ops.scenenet_fwd(x32, K, torch.float64, nnz=nnz, mode=SN_PATH_SPARSE if expect_sparse else SN_PATH_DENSE)
    assert torch.equal(pa, pe)
    assert torch.equal(ops.scenenet_fwd(x32, K, torch.float64), ops.scenenet_fwd(x32, K, torch.float64, mode=SN_PATH_DENSE))


def test_sparse_forward_full_size_agrees_with_dense():
    """config-2 full size (B = 32, 64^3)"""
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_PATH_DENSE, SN_PATH_SPARSE
    x, K = _inputs(32, (64, 64, 64), (9, 5, 5), 0.016, seed=1234, binary=True)
    ps = ops.scenenet_fwd(x, K, torch.float64, mode=SN_PATH_SPARSE)
    pd = ops.scenenet_fwd(x, K, torch.float64, mode=SN_PATH_DENSE)
    assert float((ps - pd).abs().max()) <= 2e-6
    assert float(ps.min()) >= 0.0 and float(ps.max()) < 1.0
