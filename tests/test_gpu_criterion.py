"""GPU parity of the fused criterion (csrc/criterion.cu behind the reference's GENEO_Tversky_Loss / WeightedMSE /
FocalTverskyLoss API) against the oracle's restatement of the reference criterion (oracle/model_oracle.py, pinned
to the reference's own loss values and gradients through tests/golden/ref_model.npz)."""
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"
HIST = (mo.HIST_FREQS, mo.HIST_RANGES)
KW = dict(convex_weight=5, tversky_alpha=2, tversky_beta=1, focal_gamma=4, tversky_smooth=1e-6)


def _sb():
    import scenenet_b200 as sb
    return sb


def _params(vals, frozen=()):
    d = torch.nn.ParameterDict()
    for k, v in vals.items():
        d[k] = torch.nn.Parameter(torch.tensor(float(v), dtype=torch.float32, device=DEV), requires_grad=k not in frozen)
    return d


def _data(n_shape, seed, binary, dtype=torch.float64, p_gt=0.01):
    g = torch.Generator().manual_seed(seed)
    pred = torch.relu(torch.tanh(torch.randn(n_shape, generator=g, dtype=torch.float64) * 0.5))
    if binary:
        y = (torch.rand(n_shape, generator=g) < p_gt).to(torch.float64)
    else:
        y = torch.rand(n_shape, generator=g, dtype=torch.float64) * (torch.rand(n_shape, generator=g) < 0.3)
    return pred.to(dtype).to(DEV), y.to(dtype).to(DEV)


@pytest.mark.parametrize("shape,binary,dtype", [
    ((2, 1, 32, 32, 32), True, torch.float64),
    ((3, 1, 17, 13, 11), False, torch.float64),   # ragged size (tail path), targets anywhere in [0, 1]
    ((1, 1, 16, 16, 16), False, torch.float32),
    ((4, 1, 64, 64, 64), True, torch.float64),
])
def test_fused_criterion_matches_oracle(shape, binary, dtype):
    sb = _sb()
    pred, y = _data(shape, seed=3, binary=binary, dtype=dtype)
    lambdas_v = {"lambda_cone_0": 0.3, "lambda_cy_0": 0.45, "lambda_neg_0": -0.25}   # one negative coefficient
    geneo_v = {"cy_0_radius": 2.5, "cy_0_sigma": -1.8, "cone_0_apex": 4.0, "neg_0_sigma": -0.01}
    # ours
    lam, gp = _params(lambdas_v, frozen=("lambda_cy_0",)), _params(geneo_v, frozen=("cone_0_apex",))
    p1 = pred.clone().requires_grad_(True)
    crit = sb.GENEO_Tversky_Loss(hist=HIST, **KW)
    loss = crit(p1, y, lam, gp)
    loss.backward()
    # oracle (same ops as the reference, on the same device)
    lam2, gp2 = _params(lambdas_v, frozen=("lambda_cy_0",)), _params(geneo_v, frozen=("cone_0_apex",))
    p2 = pred.clone().requires_grad_(True)
    ref = mo.geneo_tversky_criterion(p2, y, lam2, "lambda_cy_0", list(gp2.values()))
    ref.backward()
    rtol = 1e-6 if dtype == torch.float64 else 1e-4
    assert loss.dtype == ref.dtype
    assert abs(float(loss) - float(ref)) <= rtol * abs(float(ref)), (float(loss), float(ref))
    scale = float(p2.grad.abs().max())
    assert torch.allclose(p1.grad, p2.grad, rtol=rtol, atol=rtol * scale)
    for k in lambdas_v:
        a, b = lam[k].grad, lam2[k].grad
        assert (a is None) == (b is None), k
        if a is not None:
            assert float(a) == float(b), (k, float(a), float(b))
    for k in geneo_v:
        a, b = gp[k].grad, gp2[k].grad
        assert (a is None) == (b is None), k
        if a is not None:
            assert float(a) == float(b), (k, float(a), float(b))


def test_last_coefficient_penalty_active():
    """sum of the free coefficients > 1: relu(-(1 - sum + last)) is active and pushes every free coefficient down"""
    sb = _sb()
    lambdas_v = {"lambda_cone_0": 0.9, "lambda_cy_0": 0.1, "lambda_neg_0": 0.6}
    lam = _params(lambdas_v, frozen=("lambda_cy_0",))
    lam2 = _params(lambdas_v, frozen=("lambda_cy_0",))
    crit = sb.GENEO_Loss(hist=HIST, convex_weight=5)
    a = crit.cvx_loss(lam)
    a.backward()
    b = 5 * (sum(torch.relu(-v) for k, v in lam2.items() if k != "lambda_cy_0")
             + torch.relu(-(1 - sum(lam2.values()) + lam2["lambda_cy_0"])))
    b.backward()
    assert float(a) == float(b) and float(a) > 0
    for k in lambdas_v:
        assert (lam[k].grad is None) == (lam2[k].grad is None)
        if lam[k].grad is not None:
            assert float(lam[k].grad) == float(lam2[k].grad)
    gp = _params({"a": -1.0, "b": 2.0})
    r = crit.positive_regularizer(gp)
    r.backward()
    assert float(r) == 5.0 and float(gp["a"].grad) == -5.0 and float(gp["b"].grad) == 0.0
    assert crit.cvx_loss(torch.nn.ParameterDict()) == 0 and crit.positive_regularizer(torch.nn.ParameterDict()) == 0


def test_separate_terms_and_g0_variant():
    sb = _sb()
    from scenenet_b200 import ops
    pred, y = _data((2, 1, 24, 24, 24), seed=9, binary=True)
    # WeightedMSE alone and FocalTverskyLoss alone add up to the fused value
    wm = sb.WeightedMSE(hist=HIST)(pred, y)
    ft = sb.FocalTverskyLoss(2, 1, 4, 1e-6)(pred, y)
    crit = sb.GENEO_Tversky_Loss(hist=HIST, **KW)
    both = crit(pred, y, torch.nn.ParameterDict(), torch.nn.ParameterDict())
    assert abs(float(wm) + float(ft) - float(both)) <= 1e-12 * abs(float(both))
    w = mo.weight_target(y)
    assert abs(float(wm) - float(torch.mean(w * (y - pred) ** 2))) <= 1e-6 * float(wm)
    tv = sb.TverskyLoss(2, 1, 1e-6)(pred, y)
    assert abs((1 - float(tv)) - (1 - float(ft) ** 0.25)) <= 1e-9
    # G0 emitted directly by the criterion backward == g0 pass applied to its dL/dpred
    spec = crit.fused_spec()
    loss, coef, p, t = ops.criterion_fwd(pred, y, spec)
    dpred = ops.criterion_bwd(p, t, coef, spec)
    g0a = ops.criterion_bwd(p, t, coef, spec, as_g0=True)
    assert torch.equal(g0a, ops.g0(p, dpred))
    # deterministic
    loss2, coef2, _, _ = ops.criterion_fwd(pred, y, spec)
    assert torch.equal(loss, loss2) and torch.equal(coef[:spec_len(spec)], coef2[:spec_len(spec)])


def spec_len(spec):
    return len(spec.ranges)


def test_weight_table_matches_reference_ops():
    """the 10-entry table equals the reference's per-voxel get_weight_target (before the division by the mean)"""
    sb = _sb()
    crit = sb.WeightedMSE(hist=HIST, weight_alpha=1, weight_epsilon=0.1)
    ranges, w_raw = crit._weight_table()
    y = torch.tensor(ranges, dtype=torch.float64, device=DEV)
    dens = crit.get_dens_target(y)
    w = torch.max(1 - dens, torch.full_like(dens, 0.1))
    assert [float(v) for v in w] == w_raw


def test_training_loss_single_node_equals_two_step_path():
    """criterion.training_loss(model, x, y) (observer + criterion as one autograd node, G0 emitted by the criterion's
    backward) gives the bits of criterion(model(x), y, ...)"""
    sb = _sb()
    x, y = mo.synthetic_grids(2, (32, 32, 32), seed=31)
    x, y = x.to(DEV), y.to(DEV)

    def make():
        torch.manual_seed(0)
        m = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5)).to(DEV)
        with torch.no_grad():
            for name, layer in m.geneos.items():
                for pn, p in layer.geneo_params.items():
                    p.fill_(float(mo.KAT_PARAMS[f"{name}.{pn}"]))
            for ln, p in m.lambdas_dict.items():
                p.fill_(float(mo.KAT_LAMBDAS[ln]))
                p.requires_grad_(ln != mo.KAT_LAST)
        m.last_lambda = mo.KAT_LAST
        return m

    crit = sb.GENEO_Tversky_Loss(hist=HIST, **KW)
    m1 = make()
    pred1 = m1(x)
    l1 = crit(pred1, y, m1.get_cvx_coefficients(), m1.get_geneo_params())
    l1.backward()
    m2 = make()
    l2, pred2 = crit.training_loss(m2, x, y)
    l2.backward()
    assert torch.equal(pred1.detach(), pred2) and not pred2.requires_grad
    assert float(l1) == float(l2)
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert (p1.grad is None) == (p2.grad is None), n1
        if p1.grad is not None:
            assert torch.equal(p1.grad, p2.grad), (n1, float(p1.grad), float(p2.grad))
