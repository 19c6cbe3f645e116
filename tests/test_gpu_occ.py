"""GPU: the grid state buffer of sn_grid_prepare (count + one occupancy bit per voxel, ABI v3) and the mask-driven
occupancy forward (fwd_occ_kernel in csrc/stencil_fwd_sparse.cu) against the dense stencil, the scanning variant and a
float64 torch convolution."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from scenenet_b200 import ops
    return ops


def _bits(state, n):
    w = state[8:].view(torch.int32)[: (n + 31) // 32].cpu().numpy().astype("uint32")
    return np.unpackbits(w.view("uint8"), bitorder="little")[:n]


def _same_conv(x, K):
    kz, kx, ky = K.shape
    pad = []
    for k in (ky, kx, kz):  # F.pad wants the last dimension first
        pad += [(k - 1) // 2, k - 1 - (k - 1) // 2]
    return F.conv3d(F.pad(x.double(), pad), K.double()[None, None])


@pytest.mark.parametrize("shape", [(1, 1, 1, 1, 1), (2, 1, 7, 9, 13), (1, 1, 5, 3, 33), (3, 1, 20, 33, 70), (1, 1, 33, 31, 100),
                                   (2, 1, 64, 64, 64), (1, 1, 3, 5, 257)])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32, torch.uint8])
def test_prepare_state_bits(shape, dt):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    for dens in (0.0, 0.03, 0.6, 1.0):
        occ = torch.rand(shape, generator=g, device=DEV) < dens
        x = occ.to(torch.uint8) if dt == torch.uint8 else (occ * (0.25 + torch.rand(shape, generator=g, device=DEV))).to(dt)
        x32, st = ops.prepare(x)
        n = x.numel()
        ref = (x != 0).flatten().cpu().numpy().astype("uint8")
        assert int(st[0]) == int(ref.sum()) and int(st[1]) == 0
        assert np.array_equal(_bits(st, n), ref)
        words = np.add.reduceat(np.pad(ref, (0, (-n) % 32)).astype(np.int64), np.arange(0, n + (-n) % 32, 32))
        assert int(st[2]) == int((words >= 8).sum())  # clustering statistic: 32-voxel words with >= 8 voxels occupied
        assert torch.equal(x32, x.float())


@pytest.mark.parametrize("shape,ks", [((3, 1, 20, 33, 70), (9, 5, 5)), ((2, 1, 64, 64, 64), (9, 5, 5)), ((2, 1, 24, 40, 128), (9, 7, 7)),
                                      ((1, 1, 33, 31, 100), (4, 6, 5)), ((1, 1, 40, 40, 40), (9, 9, 9)), ((2, 1, 9, 17, 31), (6, 5, 5)),
                                      ((1, 1, 64, 64, 256), (9, 5, 5)), ((1, 1, 48, 48, 48), (11, 11, 11))])
@pytest.mark.parametrize("odt", [torch.float64, torch.float32])
def test_mask_driven_forward(shape, ks, odt):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(sum(shape) + sum(ks))
    for dens in (0.0, 0.02, 0.3):
        x = ((torch.rand(shape, generator=g, device=DEV) < dens) * torch.rand(shape, generator=g, device=DEV)).to(torch.float64)
        K = torch.randn(ks, generator=g, device=DEV) * 0.2
        x32, st = ops.prepare(x)
        dense = ops.scenenet_fwd(x32, K, odt, mode=1)
        mask = ops.scenenet_fwd(x32, K, odt, nnz=st, mode=2)
        scan = ops.scenenet_fwd(x32, K, odt, mode=2)
        want = torch.relu(torch.tanh(_same_conv(x32, K)))
        for name, got in (("dense", dense), ("mask", mask), ("scan", scan)):
            err = float((got.double() - want).abs().max())
            assert err < 6e-6, (name, dens, err)  # float32 accumulation of up to 1331 taps
        # capacity overflow rounds (rows with more than 128 non-zeros) are exercised by dens = 0.3


def test_float64_tanh_of_the_occupancy_forward():
    """identity kernel: pred = tanh(x) for x > 0, evaluated in float64 by tanh_pos_f64 (abs error < 1e-11)"""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(5)
    scale = torch.tensor([1e-4, 1e-2, 1.0, 12.0, 25.0, 0.3, 3.0, 0.05], device=DEV).repeat(8)
    x = (torch.rand((2, 1, 16, 16, 64), generator=g, device=DEV) * scale).float()
    K = torch.zeros((3, 3, 3), device=DEV)
    K[1, 1, 1] = 1.0
    x32, st = ops.prepare(x)
    pred = ops.scenenet_fwd(x32, K, torch.float64, nnz=st, mode=2)
    assert float((pred - torch.tanh(x.double())).abs().max()) < 1e-11


def test_per_tile_choice_between_scatter_and_stencil():
    """ABI v4: with the state buffer the choice between the occupancy-driven scatter and the dense stencil is made PER TILE
    on the device: a grid that is sparse overall (2 %) but has one dense layer gets the stencil for the tiles of that layer
    and the scatter (or a zero fill) for the rest.  Every voxel equals, bit for bit, what one of the two forced runs gives;
    the hand-off counters are back at zero afterwards (the same state serves the next forward)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(9)
    shape, ks = (4, 1, 64, 64, 64), (9, 5, 5)
    uniform = (torch.rand(shape, generator=g, device=DEV) < 0.02).double()
    layered = torch.zeros(shape, dtype=torch.float64, device=DEV)
    layered[:, :, 28:32] = (torch.rand((4, 1, 4, 64, 64), generator=g, device=DEV) < 0.7).double()  # 4.4 % overall
    mixed = uniform.clone()
    mixed[:2, :, 8:24, 8:40, :] = (torch.rand((2, 1, 16, 32, 64), generator=g, device=DEV) < 0.5).double()
    K = torch.randn(ks, generator=g, device=DEV) * 0.2
    for name, x in (("uniform", uniform), ("layered", layered), ("mixed", mixed), ("empty", torch.zeros_like(uniform))):
        x32, st = ops.prepare(x)
        dense = ops.scenenet_fwd(x32, K, torch.float64, nnz=st, mode=1)
        sparse = ops.scenenet_fwd(x32, K, torch.float64, nnz=st, mode=2)
        assert float((dense - sparse).abs().max()) < 5e-6
        for _ in range(2):  # twice on the same state buffer
            auto = ops.scenenet_fwd(x32, K, torch.float64, nnz=st)
            assert bool(((auto == dense) | (auto == sparse)).all()), name
            assert int(st[3]) == 0 and int(st[5]) == 0 and int(st[6]) == 0, name
        # tiles: 8 x 8 x 64 voxels; a tile whose halo box is more than 10 % occupied must come from the stencil
        if name == "uniform":
            assert torch.equal(auto, sparse)
        if name in ("layered", "mixed"):
            tiles = (auto != sparse).view(4, 8, 8, 8, 8, 64).any(dim=5).any(dim=4).any(dim=2)  # [b, tz, tx]
            assert int(tiles.sum()) > 0, "no tile was handed to the dense stencil"
            heavy = x.view(4, 8, 8, 8, 8, 64).sum(dim=(2, 4, 5)) > 0.3 * 8 * 8 * 64   # tiles more than 30 % occupied themselves
            assert bool((auto.view(4, 8, 8, 8, 8, 64).permute(0, 1, 3, 2, 4, 5)[heavy] ==
                         dense.view(4, 8, 8, 8, 8, 64).permute(0, 1, 3, 2, 4, 5)[heavy]).all())


@pytest.mark.parametrize("case", ["uniform", "clustered", "dense", "float32", "overflow"])
def test_quantile_model_one_forward_for_all_observers(case):
    """SCENENetQuantile inference (SCENE_Net.py:409-415) through sn_scenenet_fwd_multi against the per-observer path
    (taken when gradients are required): identical values; and the C entry point against single-observer calls"""
    import scenenet_b200 as sb
    ops = _ops()
    torch.manual_seed(3)
    qnet = sb.SCENENetQuantile({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), qs=torch.tensor([0.1, 0.5, 0.9]), device=torch.device(DEV))
    g = torch.Generator(device=DEV).manual_seed(4)
    shape = (2, 1, 32, 32, 64)
    if case == "clustered":
        x = torch.zeros(shape, dtype=torch.float64, device=DEV)
        x[:, :, 10:12] = (torch.rand((2, 1, 2, 32, 64), generator=g, device=DEV) < 0.5).double()
    elif case == "overflow":  # sparse overall, but rows with more than 128 non-zeros: several list rounds per tile
        x = torch.zeros(shape, dtype=torch.float64, device=DEV)
        x[:, :, 10, 4:8] = 1.0
        x[:, :, 20, 12:14, ::2] = 1.0
    else:
        x = (torch.rand(shape, generator=g, device=DEV) < (0.3 if case == "dense" else 0.02)).double()
    if case == "float32":
        x = x.float()
    qnet.per_observer_forward = True        # the reference's loop: one observer after the other
    per_net = qnet(x)
    per_net.square().sum().backward()
    g_ref = [None if p.grad is None else p.grad.clone() for p in qnet.parameters()]
    for p in qnet.parameters():
        p.grad = None
    qnet.per_observer_forward = False       # one forward launch for all observers, one autograd node
    fused = qnet(x)
    assert fused.shape == per_net.shape == (2, 3, 32, 32, 64) and fused.dtype == torch.float32
    assert torch.equal(fused.detach(), per_net.detach())
    fused.square().sum().backward()         # training through the fused forward: identical gradients
    for p, r in zip(qnet.parameters(), g_ref):
        assert (p.grad is None) == (r is None) and (r is None or torch.equal(p.grad, r))
    with torch.no_grad():
        assert torch.equal(qnet(x), fused.detach())
    # the entry point itself, every mode, against single-observer launches
    x32, st = ops.prepare(x)
    Ks = torch.randn((3, 9, 5, 5), generator=g, device=DEV) * 0.2
    for mode in (0, 1, 2):
        multi = ops.scenenet_fwd_multi(x32, Ks, torch.float64, nnz=st, mode=mode)
        for q in range(3):
            assert torch.equal(multi[q], ops.scenenet_fwd(x32, Ks[q], torch.float64, nnz=st, mode=mode)), (case, mode, q)
    no_state = ops.scenenet_fwd_multi(x32, Ks, torch.float32, mode=2)  # no state buffer: scanning kernel, one launch each
    for q in range(3):
        assert torch.equal(no_state[q], ops.scenenet_fwd(x32, Ks[q], torch.float32, mode=2))
