"""Pins oracle/model_oracle.py to outputs of the REAL reference (tests/golden, made by
oracle/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

NEED = {
    "cylinderv2": ["radius", "sigma"], "cylinder_kernel": ["radius", "sigma"],
    "arrow": ["apex", "cone_inc", "cone_radius", "radius", "sigma"],
    "cone_kernel": ["apex", "cone_inc", "cone_radius", "radius", "sigma"],
    "negSpherev2": ["neg_factor", "radius", "sigma"], "neg_sphere_kernel": ["neg_factor", "radius", "sigma"],
}
PARAM_SETS = {
    "kat": dict(radius=2.5, sigma=1.8, apex=4.0, cone_inc=0.3, cone_radius=2.0, neg_factor=0.2),
    "ckpt": dict(radius=1.5, sigma=0.955910, apex=0.0, cone_inc=0.565547, cone_radius=4.000988, neg_factor=0.127053),
    "wide": dict(radius=3.0, sigma=0.6, apex=7.0, cone_inc=0.12, cone_radius=1.5, neg_factor=0.9),
    "apexfull": dict(radius=0.5, sigma=1.0, apex=9.0, cone_inc=0.45, cone_radius=0.5, neg_factor=0.5),
}


def kernel_cases(golden_dir):
    ker = np.load(os.path.join(golden_dir, "ref_kernels.npz"))
    kg = np.load(os.path.join(golden_dir, "ref_kernel_grads.npz"))
    for key in ker.files:
        cname, sname, sz = key.split("|")
        ks = tuple(int(v) for v in sz.split("x"))
        yield key, cname, PARAM_SETS[sname], ks, ker[key], int(kg[key + "|seed"]), kg[key + "|g"]


def test_kernel_synthesis_matches_reference(golden_dir):
    n = 0
    for key, cname, ps, ks, Kref, seed, gref in kernel_cases(golden_dir):
        p = {k: torch.tensor(float(ps[k]), dtype=torch.float32, requires_grad=(k != "apex")) for k in NEED[cname]}
        K = mo.SYNTH[cname](p, ks)
        assert K.shape == Kref.shape, key
        scale = max(np.abs(Kref).max(), 1e-30)
        assert np.abs(K.detach().numpy() - Kref).max() <= 1e-6 * scale, key
        R = np.random.default_rng(seed).standard_normal(Kref.shape)
        (K.to(torch.float64) * torch.from_numpy(R)).sum().backward()
        g = np.array([0.0 if (k == "apex" or p[k].grad is None) else float(p[k].grad) for k in NEED[cname]])
        assert np.allclose(g, gref, rtol=2e-5, atol=1e-6 * np.abs(gref).max() + 1e-12), (key, g, gref)
        n += 1
    assert n > 150


def _load_case(gold, tag):
    names = [str(s) for s in gold[f"{tag}|grads_names"]]
    return dict(zip(names, gold[f"{tag}|grads"]))


def _check_grads(got, ref, rtol):
    for name, r in ref.items():
        g = got[name]
        if np.isnan(r):
            assert g is None, name
        else:
            assert g is not None, name
            assert abs(g - r) <= rtol * abs(r) + 1e-12, (name, g, r)


def _x575(golden_dir):
    v = np.load(os.path.join(golden_dir, "vox_sample_575.npz"))
    x = np.zeros(64 ** 3)
    x[v["restated_density_idx"]] = 1.0
    y = np.zeros(64 ** 3)
    y[v["ref_frac_idx"]] = 1.0
    return (torch.from_numpy(x).view(1, 1, 64, 64, 64), torch.from_numpy(y).view(1, 1, 64, 64, 64))


@pytest.mark.parametrize("tag,v1,last", [("kat575", False, mo.KAT_LAST), ("v1_575", True, "lambda_cone_0")])
def test_config1_kat(golden_dir, tag, v1, last):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    x, y = _x575(golden_dir)
    m = mo.OracleSceneNet(mo.KAT_GENEO_NUM, (9, 5, 5), mo.KAT_PARAMS, mo.KAT_LAMBDAS, last, v1=v1)
    pred, loss, grads = mo.fwd_bwd(m, x, y)
    p = pred.numpy().reshape(-1)
    ref = np.zeros_like(p)
    ref[gold[f"{tag}|pred_idx"]] = gold[f"{tag}|pred_val"]
    assert np.allclose(p, ref, rtol=1e-6, atol=1e-9)
    assert abs(loss - float(gold[f"{tag}|loss"])) <= 1e-9 * abs(loss)
    assert int((p >= 0.65).sum()) == int(gold[f"{tag}|pred_ge065"])
    _check_grads(grads, _load_case(gold, tag), 5e-6)


def test_checkpoint_params(golden_dir):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    cfg = json.loads(str(gold["ckpt|params"]))
    x, y = _x575(golden_dir)
    m = mo.OracleSceneNet(mo.KAT_GENEO_NUM, (9, 5, 5), cfg["params"], cfg["lambdas"], cfg["last"])
    pred, loss, grads = mo.fwd_bwd(m, x, y)
    assert abs(loss - float(gold["ckpt575|loss"])) <= 1e-9 * abs(loss)
    _check_grads(grads, _load_case(gold, "ckpt575"), 5e-6)


@pytest.mark.parametrize("ks", [(9, 7, 7), (6, 5, 5), (9, 6, 6), (7, 7, 7)])
def test_synthetic_dpred(golden_dir, ks):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    tag = f"syn32_{ks[0]}x{ks[1]}x{ks[2]}"
    x, _ = mo.synthetic_grids(2, (32, 32, 32), seed=1234)
    dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
    m = mo.OracleSceneNet(mo.KAT_GENEO_NUM, ks, mo.KAT_PARAMS, mo.KAT_LAMBDAS, "lambda_neg_0")
    pred, _, grads = mo.fwd_bwd(m, x, None, dpred)
    p = pred.numpy().reshape(-1)
    ref = np.zeros_like(p)
    ref[gold[f"{tag}|pred_idx"]] = gold[f"{tag}|pred_val"]
    assert np.allclose(p, ref, rtol=1e-6, atol=1e-9)
    _check_grads(grads, _load_case(gold, tag), 5e-6)


def test_config2_shape_b2(golden_dir):
    gold = np.load(os.path.join(golden_dir, "ref_model.npz"))
    x, y = mo.synthetic_grids(2, (64, 64, 64), seed=1234)
    m = mo.kat_model()
    pred, loss, grads = mo.fwd_bwd(m, x, y)
    assert abs(float(pred.sum()) - float(gold["syn64_crit|pred_sum"])) <= 1e-9 * float(pred.sum())
    assert abs(loss - float(gold["syn64_crit|loss"])) <= 1e-9 * abs(loss)
    _check_grads(grads, _load_case(gold, "syn64_crit"), 5e-6)


def test_invariants():
    """zero-sum slices, neg-sphere sum = -nf, pred in [0,1)."""
    m = mo.kat_model()
    K = m.kernels().detach()[:, 0]
    assert torch.all(K[0].sum(dim=(1, 2)).abs() < 1e-5)
    assert torch.all(K[1].sum(dim=(1, 2)).abs() < 1e-5)
    assert abs(float(K[2].sum()) + 0.2) < 1e-5
    x, _ = mo.synthetic_grids(1, (16, 16, 16))
    p = m.forward(x)
    assert float(p.min()) >= 0 and float(p.max()) < 1


@pytest.mark.reference
def test_live_reference_random_params():
    """Build container only: random parameter draws through the real reference vs the oracle."""
    from oracle import ref_shim
    ref_shim.install()
    from core.models.SCENE_Net import SceneNet
    import warnings
    warnings.filterwarnings("ignore")
    for seed in range(3):
        torch.manual_seed(seed)
        ref = SceneNet({'cy': 2, 'cone': 1, 'neg': 2}, (9, 7, 7))
        params = {}
        for name, layer in ref.geneos.items():
            for pn, p in layer.geneo_params.items():
                params[f"{name}.{pn}"] = float(p)
        lambdas = {k: float(v) for k, v in ref.lambdas_dict.items()}
        m = mo.OracleSceneNet({'cy': 2, 'cone': 1, 'neg': 2}, (9, 7, 7), params, lambdas, ref.last_lambda)
        x, _ = mo.synthetic_grids(1, (24, 24, 24), seed=seed)
        dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
        pr = ref(x)
        pr.backward(dpred)
        po, _, grads = mo.fwd_bwd(m, x, None, dpred)
        assert torch.allclose(pr.detach(), po, rtol=1e-6, atol=1e-9)
        for n, p in ref.named_parameters():
            if p.grad is None:
                assert grads[n] is None
            else:
                assert abs(grads[n] - float(p.grad)) <= 1e-5 * abs(float(p.grad)) + 1e-10, (n, grads[n], float(p.grad))


def test_fused_criterion_weight_table_matches_oracle_weights():
    """host logic of the fused criterion: the 10-entry table the kernels look up equals what the oracle's (= the
    reference's) per-voxel weight computation yields before the division by the mean (CPU only, no kernel call)"""
    import torch
    import scenenet_b200 as sb
    from oracle import model_oracle as mo
    crit = sb.WeightedMSE(hist=(mo.HIST_FREQS, mo.HIST_RANGES), weight_alpha=1, weight_epsilon=0.1)
    ranges, w_raw = crit._weight_table()
    assert len(ranges) == len(w_raw) == 10
    y = torch.tensor([0.0, 1.0, 0.26, 0.74, 0.5], dtype=torch.float64)
    w = mo.weight_target(y)                       # normalised by its mean
    bins = [int(torch.argmin(torch.abs(v - torch.tensor(ranges, dtype=torch.float64)))) for v in y]
    raw = torch.tensor([w_raw[b] for b in bins], dtype=torch.float32)
    assert torch.allclose(raw / raw.mean(), w.to(torch.float32), rtol=1e-6, atol=0)
    assert abs(w_raw[0] - 0.1) < 1e-7 and abs(w_raw[9] - 0.5304348) < 1e-6   # SURVEY App. C probe values
