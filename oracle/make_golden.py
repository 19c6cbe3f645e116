"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by RUNNING THE REAL REFERENCE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The outputs are small fixtures that travel with the repo; the GPU box never sees the
reference.  Every array stored here is an output of the reference's own code
(core/models/SCENE_Net.py, core/models/geneos/*.py, core/criterions/geneo_loss.py,
utils/voxelization.py::reg_on_voxel) except where the file name says `restated_` (the two
voxelization functions that no longer execute under numpy 2 / pandas 3 — see
oracle/voxel_oracle.py header).
"""
import json
import os
import sys
import warnings
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim, model_oracle as mo, voxel_oracle as vo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore")


def build_ref_model(SceneNetCls, geneo_num, ks, params, lambdas, last):
    torch.manual_seed(0)
    m = SceneNetCls(dict(geneo_num), tuple(ks))
    with torch.no_grad():
        for name, layer in m.geneos.items():
            for pn, p in layer.geneo_params.items():
                p.data = torch.tensor(float(params[f"{name}.{pn}"]), dtype=torch.float32)
        for ln in list(m.lambdas_dict.keys()):
            m.lambdas_dict[ln] = torch.nn.Parameter(torch.tensor(float(lambdas[ln]), dtype=torch.float32),
                                                    requires_grad=(ln != last))
    m.last_lambda = last
    return m


def ref_grads(m):
    out = {}
    for n, p in m.named_parameters():
        out[n] = None if p.grad is None else float(p.grad)
    return out


def sparse_pack(a):
    a = np.asarray(a)
    idx = np.flatnonzero(a)
    return idx.astype(np.int64), a.reshape(-1)[idx]


def main():
    ref_shim.install()
    from core.models.SCENE_Net import SceneNet, SCENE_Net
    from core.models.geneos import cylinder, arrow, neg_sphere
    from core.criterions.geneo_loss import GENEO_Tversky_Loss
    from utils import voxelization as Vox

    os.makedirs(GOLD, exist_ok=True)
    meta = {"torch": torch.__version__, "numpy": np.__version__}

    # ---------------------------------------------------------------- 1. kernel synthesis
    classes = {
        "cylinderv2": cylinder.cylinderv2, "cylinder_kernel": cylinder.cylinder_kernel,
        "arrow": arrow.arrow, "cone_kernel": arrow.cone_kernel,
        "negSpherev2": neg_sphere.negSpherev2, "neg_sphere_kernel": neg_sphere.neg_sphere_kernel,
    }
    param_sets = {
        "kat": dict(radius=2.5, sigma=1.8, apex=4.0, cone_inc=0.3, cone_radius=2.0, neg_factor=0.2),
        "ckpt": dict(radius=1.5, sigma=0.955910, apex=0.0, cone_inc=0.565547, cone_radius=4.000988, neg_factor=0.127053),
        "wide": dict(radius=3.0, sigma=0.6, apex=7.0, cone_inc=0.12, cone_radius=1.5, neg_factor=0.9),
        "apexfull": dict(radius=0.5, sigma=1.0, apex=9.0, cone_inc=0.45, cone_radius=0.5, neg_factor=0.5),
    }
    need = {
        "cylinderv2": ["radius", "sigma"], "cylinder_kernel": ["radius", "sigma"],
        "arrow": ["radius", "apex", "cone_radius", "cone_inc", "sigma"],
        "cone_kernel": ["radius", "apex", "cone_radius", "cone_inc", "sigma"],
        "negSpherev2": ["radius", "neg_factor", "sigma"], "neg_sphere_kernel": ["radius", "neg_factor", "sigma"],
    }
    sizes = [(9, 5, 5), (9, 7, 7), (6, 5, 5), (9, 6, 6), (9, 9, 9), (11, 11, 11), (7, 5, 3), (4, 6, 5)]
    kern = {}
    kgrads = {}
    for cname, cls in classes.items():
        for sname, ps in param_sets.items():
            for ks in sizes:
                if ps["apex"] > ks[0] and cname in ("arrow", "cone_kernel"):
                    continue
                kw = {k: torch.tensor(float(ps[k]), dtype=torch.float32, requires_grad=(k != "apex")) for k in need[cname]}
                if cname == "negSpherev2":
                    g = cls("g", ks, **kw)
                else:
                    g = cls("g", ks, False, **kw)
                K = g.kernel
                key = f"{cname}|{sname}|{ks[0]}x{ks[1]}x{ks[2]}"
                kern[key] = K.detach().numpy().astype(np.float32)
                # Jacobian^T probe: d<K, R>/dparams with a fixed pseudo-random R (float64, like the conv's grad)
                seed = zlib.crc32(key.encode())
                R = np.random.default_rng(seed).standard_normal(K.shape)
                (K.to(torch.float64) * torch.from_numpy(R)).sum().backward()
                kgrads[key + "|seed"] = np.array(seed, dtype=np.int64)
                kgrads[key + "|g"] = np.array([0.0 if (k == "apex" or kw[k].grad is None) else float(kw[k].grad)
                                               for k in sorted(need[cname])], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "ref_kernels.npz"), **kern)
    np.savez_compressed(os.path.join(GOLD, "ref_kernel_grads.npz"), **kgrads)
    print("kernels:", len(kern))

    # ---------------------------------------------------------------- 2. voxelization of sample_575
    npy = np.load(os.path.join(ref_shim.REF_ROOT, "data-sample", "sample_575.npy"))
    np.savez_compressed(os.path.join(GOLD, "sample_575.npz"), npy=npy)
    pts, labels = npy[:, 0:-1], npy[:, -1]
    ref_frac = Vox.reg_on_voxel(pts, labels, [15], voxelgrid_dims=(64, 64, 64))       # REAL reference code
    ref_frac128 = Vox.reg_on_voxel(pts, labels, [15], voxelgrid_dims=(128, 128, 128))
    g = vo.raw_grids(pts, labels, [15], (64, 64, 64))
    dens = vo.hist_on_voxel(pts, (64, 64, 64))
    from sklearn.preprocessing import MinMaxScaler
    sk = MinMaxScaler().fit_transform(g["count"].astype(np.float64).reshape(-1, 64)).reshape(64, 64, 64)
    assert np.array_equal(sk, dens), "normalize restatement != sklearn"
    fi, fv = sparse_pack(ref_frac)
    fi2, fv2 = sparse_pack(ref_frac128)
    ci, cv = sparse_pack(g["count"])
    mi, mv = sparse_pack(g["maxlab"])
    di, dv = sparse_pack(dens)
    np.savez_compressed(os.path.join(GOLD, "vox_sample_575.npz"),
                        ref_frac_idx=fi, ref_frac_val=fv, ref_frac128_idx=fi2, ref_frac128_val=fv2,
                        restated_count_idx=ci, restated_count_val=cv,
                        restated_maxlab_idx=mi, restated_maxlab_val=mv,
                        restated_density_idx=di, restated_density_val=dv,
                        restated_lin=g["lin"].astype(np.int32),
                        xyzmin=g["vg"]["xyzmin"], xyzmax=g["vg"]["xyzmax"],
                        edges=np.stack(g["vg"]["segments"]))
    meta["vox575"] = dict(n=int(len(pts)), occupied=int((g["count"] > 0).sum()), max_count=int(g["count"].max()),
                          tower_voxels=int((ref_frac > 0).sum()), sum_frac=float(ref_frac.sum()),
                          sum_lin=int(g["lin"].sum()))
    print("vox575:", meta["vox575"])

    # ---------------------------------------------------------------- 3. config-1 KAT (model + criterion)
    x = torch.from_numpy((dens > 0).astype(np.float64))[None, None]
    y = torch.from_numpy((ref_frac > 0).astype(np.float64))[None, None]
    hist_path = os.path.join(ref_shim.REF_ROOT, "core", "criterions", "hist_estimation.pickle")
    freqs, ranges = ref_shim.load_hist_pickle_cpu(hist_path)
    assert freqs.tolist() == mo.HIST_FREQS and np.allclose(ranges.numpy(), mo.HIST_RANGES)
    import core.criterions.w_mse as wm
    wm.load_pickle = ref_shim.load_hist_pickle_cpu
    crit = GENEO_Tversky_Loss(None, hist_path, 1, 0.1, 1, 5, tversky_alpha=2, tversky_beta=1, focal_gamma=4,
                              tversky_smooth=1e-6)

    torch.manual_seed(0)
    probe = SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5))
    meta["seed0_last_lambda"] = probe.last_lambda

    def run_case(tag, Cls, ks, params, lambdas, last, x, y=None, dpred=None, store_pred=True):
        m = build_ref_model(Cls, mo.KAT_GENEO_NUM, ks, params, lambdas, last)
        pred = m(x)
        if y is not None:
            loss = crit(pred, y, m.get_cvx_coefficients(), m.get_geneo_params())
            loss.backward()
        else:
            loss = None
            pred.backward(dpred)
        out = {f"{tag}|grads_names": np.array(list(ref_grads(m).keys())),
               f"{tag}|grads": np.array([np.nan if v is None else v for v in ref_grads(m).values()])}
        p = pred.detach().numpy()
        if store_pred:
            pi, pv = sparse_pack(p)
            out[f"{tag}|pred_idx"], out[f"{tag}|pred_val"] = pi, pv
        out[f"{tag}|pred_sum"] = np.array(p.sum())
        out[f"{tag}|pred_nnz"] = np.array((p > 0).sum())
        out[f"{tag}|pred_ge065"] = np.array((p >= 0.65).sum())
        if loss is not None:
            out[f"{tag}|loss"] = np.array(float(loss))
        out[f"{tag}|kernels"] = torch.stack([m.geneos[g].compute_kernel() for g in m.geneos]).detach().numpy()
        print(tag, "pred_sum", p.sum(), "loss", None if loss is None else float(loss))
        return out

    gold = {}
    gold.update(run_case("kat575", SceneNet, (9, 5, 5), mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, x, y))
    # checkpoint parameter vector (FBetaScore.ckpt), last = lambda_neg_0, apex = 0
    ck_params = {"cy_0.radius": 0.998896, "cy_0.sigma": 1.199054, "cone_0.apex": 0.0, "cone_0.cone_inc": 0.565547,
                 "cone_0.cone_radius": 4.000988, "cone_0.radius": 1.5, "cone_0.sigma": 0.955910,
                 "neg_0.neg_factor": 0.127053, "neg_0.radius": 3.000918, "neg_0.sigma": 0.605097}
    ck_lam = {"lambda_cone_0": 0.608911, "lambda_cy_0": 0.024178, "lambda_neg_0": 0.366911}
    gold.update(run_case("ckpt575", SceneNet, (9, 5, 5), ck_params, ck_lam, "lambda_neg_0", x, y))
    gold["ckpt|params"] = np.array(json.dumps({"params": ck_params, "lambdas": ck_lam, "last": "lambda_neg_0"}))
    # v1 model (SCENE_Net) on the same grid
    gold.update(run_case("v1_575", SCENE_Net, (9, 5, 5), mo.KAT_PARAMS, mo.KAT_LAMBDAS, "lambda_cone_0", x, y))
    # even extents / non-default sizes, small synthetic batch, fixed upstream gradient
    for ks in [(9, 7, 7), (6, 5, 5), (9, 6, 6), (7, 7, 7)]:
        xs, ys = mo.synthetic_grids(2, (32, 32, 32), seed=1234)
        gd = torch.Generator().manual_seed(1235)
        dpred = torch.randn(xs.shape, generator=gd, dtype=torch.float64)
        tag = f"syn32_{ks[0]}x{ks[1]}x{ks[2]}"
        gold.update(run_case(tag, SceneNet, ks, mo.KAT_PARAMS, mo.KAT_LAMBDAS, "lambda_neg_0", xs, None, dpred))
    # config-2 shape at B=2 (64^3), criterion-driven and dpred-driven
    xs, ys = mo.synthetic_grids(2, (64, 64, 64), seed=1234)
    gold.update(run_case("syn64_crit", SceneNet, (9, 5, 5), mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, xs, ys,
                         store_pred=False))
    gd = torch.Generator().manual_seed(1235)
    dpred = torch.randn(xs.shape, generator=gd, dtype=torch.float64)
    gold.update(run_case("syn64_dpred", SceneNet, (9, 5, 5), mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST, xs, None,
                         dpred, store_pred=False))
    np.savez_compressed(os.path.join(GOLD, "ref_model.npz"), **gold)
    with open(os.path.join(GOLD, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("done", meta)


if __name__ == "__main__":
    main()
