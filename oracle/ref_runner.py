"""TEST INFRASTRUCTURE ONLY — drives the reference's OWN classes (model, criterion, Lightning wrapper, checkpoints).

The same functions run (i) the unmodified reference end to end and (ii) the reference's unchanged criterion /
`LitSceneNet` on top of the CUDA drop-in model (`scenenet_cls=scenenet_b200.SceneNet`), which is the import swap of
INTEGRATION.md §A done in memory.  The reference pins its kernels to CUDA whenever a GPU is visible
(`core/models/geneos/*.py: self.device = 'cuda' if torch.cuda.is_available()`), so its CPU path on a GPU box only
runs in a process that cannot see the GPU: `run_cpu_subprocess` starts `python -m oracle.ref_runner` with
CUDA_VISIBLE_DEVICES="" and exchanges .npz / .json files.

Reference lines exercised: core/models/SCENE_Net.py:322-339 (forward), core/criterions/geneo_loss.py:145-166
(GENEO_Tversky_Loss.forward), core/lit_modules/lit_model_wrappers.py:151-174 (LitSceneNet.__init__ / training_step),
:136-148 (optimizers), experiments/.../checkpoints/*.ckpt (state_dict under the `model.` prefix).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _grads(model):
    return {n: (None if p.grad is None else float(p.grad)) for n, p in model.named_parameters()}


def _values(model):
    return {n: float(p.detach()) for n, p in model.named_parameters()}


def criterion_step(x, y, geneo_num, ks, params, lambdas, last, device="cpu", scenenet_cls=None, v1=False, dpred=None):
    """pred = model(x); loss = <reference GENEO_Tversky_Loss>(pred, y, cvx, geneo_params); loss.backward()
    (or pred.backward(dpred) when a fixed upstream gradient is given)."""
    import torch
    from oracle import ref_shim
    ref_shim.install()
    if scenenet_cls is None:
        m = ref_shim.reference_scenenet(geneo_num, ks, v1=v1)
    else:
        torch.manual_seed(0)
        m = scenenet_cls(dict(geneo_num), tuple(ks))
    m = m.to(device)
    ref_shim.set_scenenet_params(m, params, lambdas, last)
    x = x.to(device)
    pred = m(x)
    loss = None
    if dpred is not None:
        pred.backward(dpred.to(device))
    else:
        crit = ref_shim.reference_criterion(device)
        loss = crit(pred, y.to(device), m.get_cvx_coefficients(), m.get_geneo_params())
        loss.backward()
    return pred.detach(), (None if loss is None else float(loss)), _grads(m), m


def lit_training_steps(batches, geneo_num, ks, n_steps, lr=1e-3, optimizer="sgd", device="cpu", scenenet_cls=None,
                       metric_initializer=None, init=None):
    """the reference's LitSceneNet (core/lit_modules/lit_model_wrappers.py:151-174): construct under seed 0, then the loop
    Lightning runs — optimizer.zero_grad(); loss = lit.training_step(batch, i); loss.backward(); optimizer.step().
    `scenenet_cls` replaces the `SceneNet` name the wrapper imported."""
    import torch
    from oracle import ref_shim
    ref_shim.install()
    import core.lit_modules.lit_model_wrappers as lmw
    orig = lmw.SceneNet
    if scenenet_cls is not None:
        lmw.SceneNet = scenenet_cls
    try:
        torch.manual_seed(0)
        crit = ref_shim.reference_criterion(device)
        lit = lmw.LitSceneNet(dict(geneo_num), tuple(ks), crit, optimizer, lr, metric_initializer)
    finally:
        lmw.SceneNet = orig
    lit = lit.to(device)
    if init is not None:
        ref_shim.set_scenenet_params(lit.model, *init)
    opt = lit.configure_optimizers()
    losses = []
    for i in range(n_steps):
        x, y = batches[i % len(batches)]
        opt.zero_grad()
        loss = lit.training_step((x.to(device), y.to(device)), i)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses, _values(lit.model), lit


def checkpoint_forward(x, name="FBetaScore.ckpt", device="cpu", scenenet_cls=None):
    """load a shipped Lightning checkpoint into SceneNet (hyper-parameters from the checkpoint) and run forward"""
    import torch
    from oracle import ref_shim
    sd, hp = ref_shim.load_lightning_state_dict(name)
    if scenenet_cls is None:
        m = ref_shim.reference_scenenet(hp["geneo_num"], hp["kernel_size"])
    else:
        torch.manual_seed(0)
        m = scenenet_cls(dict(hp["geneo_num"]), tuple(hp["kernel_size"]))
    missing = m.load_state_dict(sd, strict=True)
    m = m.to(device)
    with torch.no_grad():
        pred = m(x.to(device))
    return pred, m, missing


# --------------------------------------------------------------------------------------------- subprocess protocol
def run_cpu_subprocess(job: dict, arrays: dict, timeout=900):
    """run `job` through the UNMODIFIED reference in a child process that cannot see the GPU; returns (meta, arrays)"""
    import numpy as np
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "in.npz"), **arrays)
        with open(os.path.join(td, "job.json"), "w") as f:
            json.dump(job, f)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONWARNINGS="ignore")
        r = subprocess.run([sys.executable, "-W", "ignore", "-m", "oracle.ref_runner", td], cwd=ROOT, env=env,
                           capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"reference subprocess failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}")
        with open(os.path.join(td, "out.json")) as f:
            meta = json.load(f)
        out = dict(np.load(os.path.join(td, "out.npz")))
    return meta, out


def _main(td):
    import numpy as np
    import torch
    assert not torch.cuda.is_available(), "the reference's CPU path needs a process without a visible GPU"
    torch.set_num_threads(os.cpu_count() or 1)
    with open(os.path.join(td, "job.json")) as f:
        job = json.load(f)
    a = np.load(os.path.join(td, "in.npz"))
    meta, out = {}, {}
    kind = job["kind"]
    if kind == "criterion_step":
        x, y = torch.from_numpy(a["x"]), torch.from_numpy(a["y"])
        dpred = torch.from_numpy(a["dpred"]) if "dpred" in a.files else None
        pred, loss, grads, m = criterion_step(x, y, job["geneo_num"], job["ks"], job["params"], job["lambdas"], job["last"],
                                              v1=job.get("v1", False), dpred=dpred)
        out["pred"] = pred.numpy()
        meta = {"loss": loss, "grads": grads}
    elif kind == "lit":
        xs, ys = a["x"], a["y"]
        batches = [(torch.from_numpy(xs[i]), torch.from_numpy(ys[i])) for i in range(xs.shape[0])]
        init = (job["params"], job["lambdas"], job["last"]) if job.get("params") else None
        losses, values, lit = lit_training_steps(batches, job["geneo_num"], job["ks"], job["n_steps"], job.get("lr", 1e-3),
                                                 job.get("optimizer", "sgd"), init=init)
        meta = {"losses": losses, "values": values, "last_lambda": lit.model.last_lambda,
                "logged": {k: float(v) for k, v in getattr(lit, "logged", {}).items()}}
        out["dummy"] = np.zeros(1)
    elif kind == "ckpt":
        pred, m, _ = checkpoint_forward(torch.from_numpy(a["x"]), job.get("name", "FBetaScore.ckpt"))
        out["pred"] = pred.numpy()
        meta = {"values": _values(m), "last_lambda": m.last_lambda}
    else:
        raise SystemExit(f"unknown job kind {kind}")
    np.savez(os.path.join(td, "out.npz"), **out)
    with open(os.path.join(td, "out.json"), "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    _main(sys.argv[1])
