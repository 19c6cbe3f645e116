"""TEST INFRASTRUCTURE ONLY — import shim for the *real* reference (/root/reference).

Only usable in the build container (the GPU box has no /root/reference).  It is
used by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/` and by the CPU tests that pin `oracle/` against the reference
when the reference tree is present.  Nothing in the product path imports this.

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

REF_ROOT = os.environ.get("SCENENET_REFERENCE", "/root/reference")


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
    "seaborn", "torchsummary", "torchviz", "pytorch_lightning",
    "pytorch_lightning.callbacks", "pytorch_lightning.loggers", "torchmetrics",
    "torchmetrics.functional", "wandb", "torchvision", "torchvision.transforms",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "core", "models"))


def install(with_pyntcloud_shim: bool = True):
    """Put the reference on sys.path with inert stubs for absent packages."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
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
