"""GPU parity at BASELINE.json's FULL sizes against the CPU ground truth (VERDICT r1 "close the tolerance gaps"):

* config 2: B = 32, 64^3, (9,5,5) — pred max-norm and each of the 11 gradients <= 1e-5 against the UNMODIFIED reference run
  on the host cores (oracle/_ref through oracle/ref_runner.py; the oracle port when the tree is absent);
* config 4: 128^3 grids, 9^3 and 15^3 kernels (B = 1) against the oracle;
* config 5: a KITTI-shaped 120 k-point scan -> (64, 64, 256) voxel grid -> SceneNet forward -> threshold 0.65, the whole chain
  against oracle voxelization + oracle model (labels bit-exact away from the threshold);
* SCENE_Net_Class.forward (SCENE_Net.py:465-466).
"""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo, ref_runner, ref_shim, voxel_oracle as vo

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-5


def _sb():
    import scenenet_b200 as sb
    return sb


def _make_model(ks, last=mo.KAT_LAST):
    torch.manual_seed(0)
    m = _sb().SceneNet(dict(mo.KAT_GENEO_NUM), tuple(ks)).to(DEV)
    return ref_shim.set_scenenet_params(m, mo.KAT_PARAMS, mo.KAT_LAMBDAS, last)


def _grads(m):
    return {n: (None if p.grad is None else float(p.grad)) for n, p in m.named_parameters()}


def _check(pred, grads, ref_pred, ref_grads, rtol=RTOL):
    ref_pred = np.asarray(ref_pred)
    err = float(np.abs(pred.detach().cpu().numpy() - ref_pred).max())
    scale = float(np.abs(ref_pred).max())
    assert err <= rtol * scale, (err, scale)
    worst = 0.0
    gmax = max(abs(v) for v in ref_grads.values() if v is not None)
    for n, r in ref_grads.items():
        if r is None:
            assert grads[n] is None, n
            continue
        e = abs(grads[n] - r)
        rel = e / max(abs(r), 1e-30)
        worst = max(worst, rel if e > 1e-7 * gmax else 0.0)
        # 1e-5 relative; entries that are tiny through cancellation: one float32 ulp of the largest gradient (test_gpu_reference.py)
        assert rel <= rtol or e <= 1e-7 * gmax, (n, grads[n], r, rel)
    return err / scale, worst


def _float64_sums(x):
    """s = sum_g lambda_g conv3d(x, K_g) in float64 on the GPU from the oracle's float32 kernels (the reference's arithmetic)"""
    import torch.nn.functional as F
    o = mo.kat_model()
    Ks = o.kernels().detach().to(DEV)
    lam = torch.stack([o.lambda_eff(n).detach().double() for n in o.geneos]).to(DEV)
    return (F.conv3d(x.to(DEV), Ks, padding="same") * lam.view(1, -1, 1, 1, 1)).sum(1, keepdim=True)


# pred = relu(tanh(s)) has a kink at s = 0 and its derivative jumps from 0 to 1 there.  Two float32 evaluations of the GENEO
# kernels differ by an ulp in some taps (ours against the reference's CPU kernels: up to 7e-8 per tap; the reference's own CPU
# and CUDA kernels differ the same way), so the SIGN of a sum with |s| < ~1e-7 — 2 of the 8.4 M voxels of the config-2 batch —
# is not defined by the model, and a voxel on the other side of the kink changes the 11 gradients by ~2e-4 relative on its
# own (measured: profiles/r2_diag_gate.log).  The full-size comparison therefore gives those voxels no upstream gradient
# (in the reference's run and in ours alike) and checks the gates everywhere else.
AMBIGUOUS = 1e-6


def test_config2_full_size_vs_reference():
    """BASELINE config 2 at its full size: the bench workload itself (seed 1234 grids, seed 1235 upstream gradient)"""
    x, _ = mo.synthetic_grids(32, (64, 64, 64), seed=1234)
    dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
    s_ref = _float64_sums(x)
    amb = ((s_ref.abs() < AMBIGUOUS) & (s_ref != 0)).cpu()
    assert 1 <= int(amb.sum()) <= 8, int(amb.sum())
    dpred[amb] = 0.0
    if ref_shim.available():
        job = dict(kind="criterion_step", geneo_num=mo.KAT_GENEO_NUM, ks=[9, 5, 5], params=mo.KAT_PARAMS, lambdas=mo.KAT_LAMBDAS, last=mo.KAT_LAST)
        meta, out = ref_runner.run_cpu_subprocess(job, dict(x=x.numpy(), y=np.zeros(1), dpred=dpred.numpy()))
        ref_pred, ref_grads, who = out["pred"], meta["grads"], "reference"
    else:
        pr, _, gr = mo.fwd_bwd(mo.kat_model(), x, None, dpred)
        ref_pred, ref_grads, who = pr.numpy(), gr, "oracle port"
    for modes in ((0, 0), (1, 1), (2, 2)):  # per-tile choice, dense stencils, occupancy-driven kernels
        m = _make_model((9, 5, 5))
        m.path_modes = modes
        pred = m(x.to(DEV))
        flips = ((pred.detach() > 0) != (s_ref > 0)) & ~amb.to(DEV)
        assert int(flips.sum()) == 0, (modes, int(flips.sum()))  # the relu gates agree wherever the model defines them
        pred.backward(dpred.to(DEV))
        e, w = _check(pred, _grads(m), ref_pred, ref_grads)
        print(f"config 2 full size vs {who}, path modes {modes}: pred max-norm rel {e:.2e}, worst grad rel {w:.2e}, "
              f"{int(amb.sum())} voxels with 0 < |s| < {AMBIGUOUS:g} excluded")


def test_config2_full_size_training_step_vs_reference():
    """config 2(ii): the whole batch through the fused drop-in criterion against the reference's own GENEO_Tversky_Loss"""
    x, y = mo.synthetic_grids(32, (64, 64, 64), seed=1234)
    if ref_shim.available():
        job = dict(kind="criterion_step", geneo_num=mo.KAT_GENEO_NUM, ks=[9, 5, 5], params=mo.KAT_PARAMS, lambdas=mo.KAT_LAMBDAS, last=mo.KAT_LAST)
        meta, out = ref_runner.run_cpu_subprocess(job, dict(x=x.numpy(), y=y.numpy()))
        ref_pred, ref_loss, ref_grads = out["pred"], meta["loss"], meta["grads"]
    else:
        pr, ref_loss, ref_grads = mo.fwd_bwd(mo.kat_model(), x, y)
        ref_pred = pr.numpy()
    sb = _sb()
    crit = sb.GENEO_Tversky_Loss(hist=(mo.HIST_FREQS, mo.HIST_RANGES), weight_alpha=1, weight_epsilon=0.1, mse_weight=1,
                                 convex_weight=5, tversky_alpha=2, tversky_beta=1, focal_gamma=4, tversky_smooth=1e-6)
    for single_node in (False, True):
        m = _make_model((9, 5, 5))
        if single_node:
            loss, pred = crit.training_loss(m, x.to(DEV), y.to(DEV))
        else:
            pred = m(x.to(DEV))
            loss = crit(pred, y.to(DEV), m.get_cvx_coefficients(), m.get_geneo_params())
        loss.backward()
        assert abs(float(loss) - ref_loss) <= RTOL * abs(ref_loss), (float(loss), ref_loss)
        e, w = _check(pred, _grads(m), ref_pred, ref_grads)
        print(f"config 2(ii) full size, single_node={single_node}: loss rel {abs(float(loss) - ref_loss) / abs(ref_loss):.2e}, worst grad rel {w:.2e}")


@pytest.mark.parametrize("k", [9, 15])
def test_config4_full_size_128_vs_oracle(k):
    """BASELINE config 4 shapes: one 128^3 grid, cubic k^3 kernels (T = 729 / 3375 taps), fwd + bwd"""
    x, _ = mo.synthetic_grids(1, (128, 128, 128), seed=4)
    dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    pr, _, gr = mo.fwd_bwd(mo.kat_model((k, k, k)), x, None, dpred)
    m = _make_model((k, k, k))
    pred = m(x.to(DEV))
    pred.backward(dpred.to(DEV))
    e, w = _check(pred, _grads(m), pr.numpy(), gr)
    print(f"config 4, 128^3, {k}^3 kernels: pred max-norm rel {e:.2e}, worst grad rel {w:.2e}")


def _kitti_scan(n=120_000, seed=0):
    """SemanticKITTI-shaped scan: range-weighted rings in +-50 m, z in [-3, 3], float32-valued
    coordinates stored as float64, labels from SemanticKITTI ids with pole = 80"""
    rng = np.random.default_rng(seed)
    r = 50.0 * np.sqrt(rng.random(n)) * rng.random(n)
    th = rng.random(n) * 2 * np.pi
    z = np.clip(rng.normal(-1.5, 0.6, n), -3, 3)
    poles = rng.random(n) < 0.01
    z[poles] = rng.uniform(-2, 3, poles.sum())
    pts = np.stack([r * np.cos(th), r * np.sin(th), z], 1).astype(np.float32).astype(np.float64)
    lab = rng.choice([40.0, 44.0, 48.0, 50.0, 70.0, 71.0, 72.0], n)
    lab[poles] = 80.0
    return pts, lab


def test_config5_chain_voxelize_infer_threshold():
    sb = _sb()
    pts, lab = _kitti_scan()
    dims = (64, 64, 256)  # (n_x, n_y, n_z) as in semKITTI.py:453-454
    out = sb.voxel_ops.voxelize_clouds(torch.from_numpy(pts).to(DEV), None, dims, torch.from_numpy(lab).to(DEV), [80],
                                       want=("occ", "occ_keep", "keep_count"), occ_dtype=torch.float64)
    g = vo.raw_grids(pts, lab, [80], dims)
    assert np.array_equal(out["count"][0].cpu().numpy(), g["count"])
    assert np.array_equal(out["keep_count"][0].cpu().numpy(), g["keep"])
    x = out["occ"][0][None, None]
    assert tuple(x.shape) == (1, 1, 256, 64, 64) and np.array_equal(x[0, 0].cpu().numpy(), (g["count"] > 0).astype(np.float64))
    m = _make_model((9, 5, 5))
    with torch.no_grad():
        pred = m(x)
    from scenenet_b200.utils import voxelization as Vox
    lab_gpu = Vox.prob_to_label(pred, 0.65).cpu().numpy()
    with torch.no_grad():
        pr = mo.kat_model().forward(x.cpu()).numpy()
    assert np.abs(pred.cpu().numpy() - pr).max() <= RTOL * np.abs(pr).max()
    safe = np.abs(pr - 0.65) > 1e-5
    assert np.array_equal(lab_gpu[safe], (pr >= 0.65).astype(pr.dtype)[safe])
    assert int((lab_gpu != (pr >= 0.65)).sum()) <= 2
    print(f"config 5 chain: {int((g['count'] > 0).sum())} occupied voxels, {int((pr >= 0.65).sum())} positive labels")


def test_scene_net_class_forward():
    """SCENE_Net_Class.forward = (gnet(x) >= tau).to(x.dtype) (SCENE_Net.py:465-466)"""
    sb = _sb()
    torch.manual_seed(1)
    clf = sb.SCENE_Net_Class({'cy': 1, 'cone': 1, 'neg': 1}, plot=False).to(DEV)
    assert 0.0 <= float(clf.get_threshold()) <= 0.4 and isinstance(clf.gnet, sb.SCENE_Net)
    x, _ = mo.synthetic_grids(2, (24, 24, 24), seed=9, p_occ=0.05)
    x = x.to(DEV)
    for tau in (None, 0.05, 0.5):
        if tau is not None:
            with torch.no_grad():
                clf.tau.fill_(tau)  # an in-place update (optimizer step) must refresh the cached host copy
        out = clf(x)
        ref = (clf.gnet(x) >= clf.tau).to(x.dtype)
        assert out.dtype == x.dtype and out.shape == x.shape and torch.equal(out, ref)
    assert set(clf.get_cvx_coefficients().keys()) == {"lambda_cy_0", "lambda_cone_0", "lambda_neg_0"}
    assert float(out.sum()) >= 0 and set(torch.unique(out).tolist()) <= {0.0, 1.0}
    xf = x.to(torch.float32)
    assert clf(xf).dtype == torch.float32
