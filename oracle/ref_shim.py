"""TEST INFRASTRUCTURE ONLY — import shim for the *real* reference.

The reference tree is `/root/reference` in the build container and the unmodified copy `oracle/_ref`
(made by `oracle/fetch_ref.py`, git-ignored, shipped by gpurun) on the GPU box.  It is
used by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/`, by the CPU tests that pin `oracle/` against the reference, by the `-m gpu`
tests that run the reference's own criterion / Lightning wrapper / checkpoints over the CUDA
modules, and by `bench.py --impl reference`.  Nothing in the product path imports this.

The reference imports a dozen plotting / IO packages at module top that are not
installed here and play no part in the arithmetic: they are
replaced by inert stub modules.  pyntcloud IS arithmetic (VoxelGrid.compute),
it is absent too, so the voxelization functions of the reference are exercised
through `oracle.voxel_oracle.PyntCloudShim` (a functional restatement of
pyntcloud 0.1.6's VoxelGrid — "parity unpinned" for that third-party step).
"""
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref_root() -> str:
    env = os.environ.get("SCENENET_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isdir(os.path.join(cand, "core", "models")):
            return cand
    return "/root/reference"


REF_ROOT = _find_ref_root()
CKPT_DIR = os.path.join(REF_ROOT, "experiments", "scenenet_ts40k", "wandb", "run-20230217_161733-bwsbqxgs", "files",
                        "checkpoints")
HIST_PICKLE = os.path.join(REF_ROOT, "core", "criterions", "hist_estimation.pickle")


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub / a do-nothing callable."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behave like a package so `import a.b` works
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        full = f"{self.__name__}.{item}"
        if full in sys.modules:
            return sys.modules[full]
        obj = _StubObj(full)
        setattr(self, item, obj)
        return obj


class _StubObj:
    def __init__(self, name="stub"):
        self._name = name

    def __call__(self, *a, **k):
        return _StubObj(self._name + "()")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _StubObj(f"{self._name}.{item}")

    def __mro_entries__(self, bases):  # allow `class X(pl.LightningModule)`
        return (object,)

    def __iter__(self):
        return iter(())


_STUBBED = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
    "mpl_toolkits", "mpl_toolkits.mplot3d",
    "sympytorch", "IPython", "IPython.display", "open3d", "laspy", "webcolors",
    "seaborn", "torchsummary", "torchviz", "pytorch_lightning.callbacks", "pytorch_lightning.loggers", "torchmetrics",
    "torchmetrics.functional", "wandb", "torchvision", "torchvision.transforms",
]


def _lightning_standin():
    """pytorch_lightning 1.9 is absent: the smallest stand-in under which the reference's
    `core/lit_modules/lit_model_wrappers.py` runs UNCHANGED — `LightningModule` is an `nn.Module` with
    `save_hyperparameters` (reads the named arguments from the caller's frame, like Lightning), `hparams`,
    `log` / `log_dict` (recorded in `self.logged`), `trainer`; `seed_everything` seeds python / numpy / torch."""
    import inspect
    import random
    import numpy as np
    import torch

    class AttributeDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    class LightningModule(torch.nn.Module):
        trainer = None
        logger = None

        @property
        def hparams(self):
            if "_hparams" not in self.__dict__:
                self.__dict__["_hparams"] = AttributeDict()
            return self.__dict__["_hparams"]

        def save_hyperparameters(self, *names, **kw):
            frame = inspect.currentframe().f_back
            loc = frame.f_locals
            for n in names:
                if n in loc:
                    self.hparams[n] = loc[n]

        def log(self, name, value, **kw):
            self.__dict__.setdefault("logged", {})[name] = value

        def log_dict(self, d, **kw):
            for k, v in d.items():
                self.log(k, v)

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class Callback:
        pass

    def seed_everything(seed=None, workers=False):
        seed = int(seed or 0)
        random.seed(seed)
        np.random.seed(seed % (2 ** 32))
        torch.manual_seed(seed)
        return seed

    mod = _Stub("pytorch_lightning")  # anything else (Trainer, loggers ...) stays an inert stub
    mod.LightningModule = LightningModule
    mod.LightningDataModule = LightningDataModule
    mod.Callback = Callback
    mod.seed_everything = seed_everything
    mod.STANDIN = True
    return mod


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "core", "models"))


def install(with_pyntcloud_shim: bool = True):
    """Put the reference on sys.path with inert stubs for absent packages."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    try:
        import pytorch_lightning  # noqa: F401  (the real one, if a box has it)
    except Exception:
        sys.modules["pytorch_lightning"] = _lightning_standin()
    for name in _STUBBED:
        try:
            if name not in sys.modules:
                importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    for name in _STUBBED:  # link children onto parents
        if "." in name and isinstance(sys.modules.get(name), _Stub):
            parent, child = name.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _Stub):
                setattr(sys.modules[parent], child, sys.modules[name])
    if with_pyntcloud_shim and "pyntcloud" not in sys.modules:
        from oracle import voxel_oracle
        mod = types.ModuleType("pyntcloud")
        mod.PyntCloud = voxel_oracle.PyntCloudShim
        sys.modules["pyntcloud"] = mod
        # open3d is only a container for float64 points in the reference
        o3d = sys.modules["open3d"]
        if isinstance(o3d, _Stub):
            o3d.geometry = types.SimpleNamespace(PointCloud=voxel_oracle.O3DPointCloudShim)
            o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: a)
    for p in (REF_ROOT, os.path.join(REF_ROOT, "scripts")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # the reference uses the removed alias np.float (torch_transforms.py:13)
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]


def load_hist_pickle_cpu(path):
    """hist_estimation.pickle stores CUDA tensors; load them on a CPU-only host."""
    import cloudpickle
    import torch
    orig = torch.storage._load_from_bytes
    import io
    torch.storage._load_from_bytes = lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
    try:
        with open(path, "rb") as f:
            return cloudpickle.load(f)
    finally:
        torch.storage._load_from_bytes = orig


# --------------------------------------------------------------------------------------
# helpers for tests / bench.py --impl reference: build the reference's OWN objects
# --------------------------------------------------------------------------------------
def set_scenenet_params(model, params: dict, lambdas: dict, last: str):
    """overwrite the 13 scalars of a (reference or mirror) SceneNet / SCENE_Net in place; `last` becomes the
    frozen, derived lambda (`SCENE_Net.py:285-291`)"""
    import torch
    with torch.no_grad():
        for name, layer in model.geneos.items():
            for pn, p in layer.geneo_params.items():
                p.fill_(float(params[f"{name}.{pn}"]))
        for ln, p in model.lambdas_dict.items():
            p.fill_(float(lambdas[ln]))
            p.requires_grad_(ln != last)
    model.last_lambda = last
    return model


def reference_scenenet(geneo_num, kernel_size, params=None, lambdas=None, last=None, v1=False, seed=0):
    """the reference's SceneNet (or v1 SCENE_Net) on the CPU, constructed under `seed`"""
    import torch
    install()
    from core.models.SCENE_Net import SceneNet, SCENE_Net
    torch.manual_seed(seed)
    m = (SCENE_Net if v1 else SceneNet)(dict(geneo_num), tuple(kernel_size))
    if params is not None:
        set_scenenet_params(m, params, lambdas, last)
    return m


def reference_criterion(device="cpu", **kw):
    """the reference's own GENEO_Tversky_Loss (core/criterions/geneo_loss.py:145-166) with the shipped histogram
    pickle and the defaults of defaults_config.yml:57-77.  The class pins its tensors to CUDA whenever a GPU is
    visible (`w_mse.py:56`); `device` moves them (attribute assignment only — the class is unchanged)."""
    import torch
    install()
    from core.criterions.geneo_loss import GENEO_Tversky_Loss
    args = dict(weight_alpha=1, weight_epsilon=0.1, mse_weight=1, convex_weight=5, tversky_alpha=2, tversky_beta=1,
                focal_gamma=4, tversky_smooth=1e-6)
    args.update(kw)
    orig = torch.storage._load_from_bytes
    if not torch.cuda.is_available():  # the pickle stores CUDA tensors
        import io
        torch.storage._load_from_bytes = lambda b: torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)
    try:
        crit = GENEO_Tversky_Loss(None, HIST_PICKLE, args.pop("weight_alpha"), args.pop("weight_epsilon"), args.pop("mse_weight"),
                                  args.pop("convex_weight"), **args)
    finally:
        torch.storage._load_from_bytes = orig
    dev = torch.device(device)
    crit.device = dev
    crit.freqs, crit.ranges = crit.freqs.to(dev), crit.ranges.to(dev)
    return crit


def load_lightning_state_dict(name="FBetaScore.ckpt"):
    """`model.`-prefixed SceneNet entries of a shipped Lightning checkpoint + its hyper-parameters"""
    import torch
    install()
    ck = torch.load(os.path.join(CKPT_DIR, name), map_location="cpu", weights_only=False)
    sd = {k[len("model."):]: v for k, v in ck["state_dict"].items() if k.startswith("model.")}
    return sd, ck.get("hyper_parameters", {})
