"""GPU parity of the two tap-gradient kernels (dense stencil, occupancy-driven) and of the device-side
selection between them, through the C ABI (sn_grid_prepare, sn_scenenet_tapgrad).

Checker: a float64 cross-correlation of the same inputs (torch on the GPU: the batch is laid out as
channels so that one conv3d call returns W[t] = sum_{b,v} G0[b,v] * xpad[b,v+t], the weight gradient of
SCENE_Net.py:325 for a single output channel).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_tapgrad(x, g0, ks):
    kz, kx, ky = ks
    pl = [(k - 1) // 2 for k in ks]
    pr = [k - 1 - l for k, l in zip(ks, pl)]
    xp = F.pad(x[:, 0].to(torch.float64), (pl[2], pr[2], pl[1], pr[1], pl[0], pr[0]))  # [B,Z+,X+,Y+]
    return F.conv3d(xp[None], g0[:, 0].to(torch.float64)[None])[0, 0]                   # [kz,kx,ky]


def _inputs(B, grid, density, seed, binary):
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand((B, 1, *grid), generator=g) < density
    val = torch.ones((B, 1, *grid)) if binary else torch.rand((B, 1, *grid), generator=g) + 0.25
    x = (occ * val).to(torch.float32)
    g0 = torch.randn((B, 1, *grid), generator=g).to(torch.float32)
    return x.to(DEV), g0.to(DEV)


CASES = [
    # B, grid (Z,X,Y), kernel, density, binary
    (2, (64, 64, 64), (9, 5, 5), 0.016, True),      # config-2 shape
    (3, (20, 17, 23), (9, 5, 5), 0.05, False),      # ragged, Y % 4 != 0 -> plain-load path
    (2, (24, 24, 24), (9, 7, 7), 0.02, True),
    (1, (16, 16, 16), (3, 3, 3), 0.3, False),       # far above the selection threshold
    (1, (12, 12, 12), (5, 5, 4), 0.1, False),       # ky = 4: no dense instantiation
    (2, (16, 16, 16), (4, 6, 5), 0.02, False),      # even extents (asymmetric 'same' padding)
    (1, (32, 32, 32), (11, 11, 11), 0.016, True),   # 6 tap chunks x 2 slices
    (1, (24, 24, 64), (15, 15, 15), 0.016, True),   # 14 tap chunks, single-stage ring
    (2, (8, 8, 128), (9, 5, 5), 0.016, True),       # several y tiles
    (1, (40, 8, 32), (1, 1, 1), 0.5, False),        # single tap
    (2, (64, 64, 64), (9, 5, 5), 0.0, True),        # empty grids
    (1, (64, 64, 64), (9, 5, 5), 1.0, False),       # full grids
]


@pytest.mark.parametrize("B,grid,ks,density,binary", CASES)
def test_sparse_and_dense_tapgrad_match_float64_reference(B, grid, ks, density, binary):
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
    x, g0 = _inputs(B, grid, density, seed=hash((B, grid, ks)) % 1000, binary=binary)
    ref = _ref_tapgrad(x, g0, ks)
    # float32 products summed in float32 within a lane for at most a few hundred terms, float64 above that
    terms = _ref_tapgrad(x.abs(), g0.abs(), ks)
    tol = 2e-6 * terms + 1e-12
    Ws = ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_SPARSE)
    Wd = ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_DENSE)
    assert Ws.shape == ref.shape and Ws.dtype == torch.float64
    assert bool(((Ws - ref).abs() <= tol).all()), f"sparse: max err {(Ws - ref).abs().max():.3e}"
    assert bool(((Wd - ref).abs() <= tol).all()), f"dense: max err {(Wd - ref).abs().max():.3e}"
    # determinism
    assert torch.equal(Ws, ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_SPARSE))


@pytest.mark.parametrize("density,expect_sparse", [(0.016, True), (0.08, True), (0.12, False), (1.0, False)])
def test_device_side_selection(density, expect_sparse):
    """AUTO + the count from sn_grid_prepare picks the kernel on the device: the result is bit-identical to
    the forced run of the expected kernel (threshold: 10 % occupancy)."""
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_TAPGRAD_AUTO, SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
    ks = (9, 5, 5)
    x, g0 = _inputs(2, (32, 32, 64), density, seed=7, binary=False)
    for dt in (torch.float64, torch.float32, torch.uint8):
        xin = (x != 0).to(dt) if dt == torch.uint8 else x.to(dt)
        x32, nnz = ops.prepare(xin)
        assert int(nnz[0]) == int((xin != 0).sum())
        assert x32.dtype == torch.float32 and torch.equal(x32, xin.to(torch.float32))
        Wa = ops.tapgrad(x32, g0, ks, nnz=nnz, mode=SN_TAPGRAD_AUTO)
        # (forced runs with the same state buffer: binary grids take their voxels from its occupancy bits, another order)
        We = ops.tapgrad(x32, g0, ks, nnz=nnz, mode=SN_TAPGRAD_SPARSE if expect_sparse else SN_TAPGRAD_DENSE)
        Wo = ops.tapgrad(x32, g0, ks, nnz=nnz, mode=SN_TAPGRAD_DENSE if expect_sparse else SN_TAPGRAD_SPARSE)
        assert torch.equal(Wa, We)
        assert torch.allclose(Wa, Wo, rtol=1e-4, atol=1e-4 * float(Wo.abs().max()))
        # the same state buffer serves a second backward (the kernels leave its counters untouched)
        assert torch.equal(ops.tapgrad(x32, g0, ks, nnz=nnz, mode=SN_TAPGRAD_AUTO), We)
    # no count -> dense
    assert torch.equal(ops.tapgrad(x, g0, ks), ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_DENSE))


def test_sparse_tapgrad_full_size_properties():
    """config-2 full size (B = 32, 64^3): linearity in G0 and agreement with the dense stencil."""
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
    ks = (9, 5, 5)
    x, g0 = _inputs(32, (64, 64, 64), 0.016, seed=1234, binary=True)
    g1 = torch.randn(g0.shape, generator=torch.Generator().manual_seed(5)).to(DEV)
    Wa = ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_SPARSE)
    Wb = ops.tapgrad(x, g1, ks, mode=SN_TAPGRAD_SPARSE)
    Wab = ops.tapgrad(x, 2.0 * g0 + g1, ks, mode=SN_TAPGRAD_SPARSE)
    scale = float(Wab.abs().max())
    assert torch.allclose(Wab, 2.0 * Wa + Wb, rtol=0, atol=2e-5 * scale)
    Wd = ops.tapgrad(x, g0, ks, mode=SN_TAPGRAD_DENSE)
    assert torch.allclose(Wa, Wd, rtol=0, atol=2e-5 * float(Wd.abs().max()))


FUSED_CASES = [
    # B, grid (Z,X,Y), kernel, density, binary, dtype
    (2, (64, 64, 64), (9, 5, 5), 0.016, True, torch.float64),     # config-2 shape
    (2, (64, 64, 64), (9, 5, 5), 0.016, False, torch.float32),    # density values, float32 predictions
    (3, (20, 17, 32), (9, 5, 5), 0.05, False, torch.float64),     # ragged z / x, 8 x 16 x 32 tiles
    (1, (32, 32, 32), (11, 11, 11), 0.016, True, torch.float64),  # 6 tap chunks
    (1, (24, 24, 64), (15, 15, 15), 0.016, True, torch.float64),  # 14 tap chunks, ring of 32 z-rows
    (2, (8, 8, 128), (9, 5, 5), 0.016, True, torch.float64),      # several y tiles, one z tile
    (1, (40, 8, 32), (1, 1, 1), 0.5, False, torch.float64),       # single tap
    (2, (16, 16, 32), (4, 6, 5), 0.02, False, torch.float64),     # even extents (asymmetric 'same' padding)
    (1, (72, 24, 64), (9, 5, 5), 0.3, False, torch.float64),      # lists longer than one window
    (2, (64, 64, 64), (9, 5, 5), 0.0, True, torch.float64),       # empty grids
]


@pytest.mark.parametrize("B,grid,ks,density,binary,dt", FUSED_CASES)
def test_bwd_entry_point_equals_g0_plus_tapgrad(B, grid, ks, density, binary, dt):
    """sn_scenenet_bwd (G0 pass + tap gradient in one call, with and without the grid state buffer) agrees with the float64
    cross-correlation of the same G0 and is bit-identical to sn_scenenet_g0 + sn_scenenet_tapgrad in every mode."""
    from scenenet_b200 import ops
    from scenenet_b200._lib import SN_TAPGRAD_AUTO, SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
    x, _ = _inputs(B, grid, density, seed=hash((B, grid, ks)) % 1000, binary=binary)
    g = torch.Generator().manual_seed(11)
    pred = torch.relu(torch.tanh(torch.randn((B, 1, *grid), generator=g, dtype=torch.float64))).to(dt).to(DEV)
    dpred = torch.randn((B, 1, *grid), generator=g, dtype=torch.float64).to(dt).to(DEV)
    x32, state = ops.prepare(x)
    g0 = ops.g0(pred, dpred)
    ref = _ref_tapgrad(x, g0, ks)
    tol = 2e-6 * _ref_tapgrad(x.abs(), g0.abs(), ks) + 1e-12
    Wf = ops.scenenet_bwd(x32, pred, dpred, ks, nnz=state, mode=SN_TAPGRAD_SPARSE)
    assert bool(((Wf - ref).abs() <= tol).all()), f"max err {(Wf - ref).abs().max():.3e}"
    # deterministic, and the state buffer serves a second backward
    assert torch.equal(Wf, ops.scenenet_bwd(x32, pred, dpred, ks, nnz=state, mode=SN_TAPGRAD_SPARSE))
    # device-side choice: the occupancy-driven kernel below the occupancy threshold, the dense stencil above it
    Wa = ops.scenenet_bwd(x32, pred, dpred, ks, nnz=state, mode=SN_TAPGRAD_AUTO)
    if density <= 0.05:
        assert torch.equal(Wa, Wf)
    elif ks[2] in (3, 5, 6, 7, 9, 11, 13, 15):
        assert torch.equal(Wa, ops.scenenet_bwd(x32, pred, dpred, ks, mode=SN_TAPGRAD_DENSE))
    else:  # widths without a dense instantiation always take the occupancy-driven kernel
        assert torch.equal(Wa, Wf)
    # without a state buffer: the two-kernel path
    assert torch.equal(ops.scenenet_bwd(x32, pred, dpred, ks, mode=SN_TAPGRAD_SPARSE), ops.tapgrad(x32, g0, ks, mode=SN_TAPGRAD_SPARSE))
    assert torch.equal(ops.tapgrad(x32, g0, ks, nnz=state, mode=SN_TAPGRAD_SPARSE), Wf)
    # binary grids with a state buffer are walked through the occupancy bits (no x tile): same sum, another order
    Wx = ops.tapgrad(x32, g0, ks, mode=SN_TAPGRAD_SPARSE)
    assert bool(((Wx - ref).abs() <= tol).all()) and (binary or torch.equal(Wx, Wf))
