"""CPU tests of the reference hand-over (`-m reference`: need the reference tree, i.e. the build container or a box
that received `oracle/_ref`): the shipped copy is byte-identical to /root/reference, and the reference's own
LitSceneNet / GENEO_Tversky_Loss / checkpoints run under the Lightning stand-in and agree with the oracle port."""
import os

import numpy as np
import pytest
import torch

from oracle import fetch_ref, model_oracle as mo, ref_runner, ref_shim

pytestmark = pytest.mark.reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shipped_reference_copy_is_unmodified():
    if not os.path.isdir("/root/reference"):
        assert fetch_ref.verify(), "oracle/_ref does not match its manifest"
        return
    manifest = fetch_ref.fetch("/root/reference")  # idempotent: copies only what is missing / changed
    assert fetch_ref.verify()
    assert "core/models/SCENE_Net.py" in manifest and "core/lit_modules/lit_model_wrappers.py" in manifest
    assert "core/criterions/hist_estimation.pickle" in manifest and "data-sample/sample_575.npy" in manifest
    assert any(k.endswith("FBetaScore.ckpt") for k in manifest)
    # never tracked by git
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "oracle/_ref/" in f.read()
    assert not os.path.exists(os.path.join(ROOT, ".gpurunignore")) or "oracle/_ref" not in open(os.path.join(ROOT, ".gpurunignore")).read()


def test_reference_step_in_gpu_less_subprocess_equals_oracle():
    x, y = mo.synthetic_grids(2, (32, 32, 32), seed=7)
    job = dict(kind="criterion_step", geneo_num=mo.KAT_GENEO_NUM, ks=[9, 5, 5], params=mo.KAT_PARAMS, lambdas=mo.KAT_LAMBDAS, last=mo.KAT_LAST)
    meta, out = ref_runner.run_cpu_subprocess(job, dict(x=x.numpy(), y=y.numpy()))
    pr, lo, gr = mo.fwd_bwd(mo.kat_model(), x, y)
    assert np.abs(out["pred"] - pr.numpy()).max() <= 1e-12
    assert abs(meta["loss"] - lo) <= 1e-12 * abs(lo)
    for n, r in gr.items():
        g = meta["grads"][n]
        assert (g is None) == (r is None), n
        if r is not None:
            assert abs(g - r) <= 2e-6 * abs(r) + 1e-10, (n, g, r)  # float32 autograd noise of the reference's synthesis


def test_lit_scenenet_runs_under_the_lightning_standin():
    x, y = mo.synthetic_grids(1, (32, 32, 32), seed=3, p_gt=2e-3)
    losses, values, lit = ref_runner.lit_training_steps([(x, y)], mo.KAT_GENEO_NUM, (9, 5, 5), 2, 1e-3, "sgd",
                                                        init=(mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST))
    assert type(lit).__name__ == "LitSceneNet" and lit.hparams["kernel_size"] == (9, 5, 5) and lit.hparams["optimizer_name"] == "sgd"
    # first step = the oracle's loss on the same inputs
    _, lo, gr = mo.fwd_bwd(mo.kat_model(), x, y)
    assert abs(losses[0] - lo) <= 1e-10 * abs(lo)
    # one SGD step moved every trainable parameter by -lr * grad
    n = "geneos.cy_0.geneo_params.sigma"
    assert losses[1] != losses[0] and abs(values[n] - mo.KAT_PARAMS["cy_0.sigma"]) > 0


def test_checkpoint_vector_matches_survey():
    sd, hp = ref_shim.load_lightning_state_dict("FBetaScore.ckpt")
    assert tuple(hp["kernel_size"]) == (9, 5, 5) and hp["geneo_num"] == {'cy': 1, 'cone': 1, 'neg': 1}
    assert abs(float(sd["geneos.cone_0.geneo_params.cone_radius"]) - 4.000988) < 1e-6
    assert abs(float(sd["lambdas_dict.lambda_cone_0"]) - 0.608911) < 1e-6
