"""scenenet_b200 — B200-native (sm_100a) implementation of SCENE-Net's data-parallel hot path:
voxelization -> GENEO kernel synthesis -> 3-D stencil + observer -> backward onto the 13 scalars.

Importing this package loads the in-tree CUDA library (libscenenet_b200.so) through its C ABI
(include/scenenet_b200.h) and FAILS if it is missing: there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is not built)
from . import ops, voxel_ops, dist  # noqa: F401
from .core.models.SCENE_Net import GENEO_Layer, SCENE_Net, SceneNet, SCENENetQuantile, SCENE_Net_Class  # noqa: F401
from .core.datasets.torch_transforms import Voxelization, ToTensor, ToFullDense  # noqa: F401
from .core.datasets.ts40k import TS40K, TS40KDeviceLoader  # noqa: F401
from .core.criterions.geneo_loss import GENEO_Loss, GENEO_Tversky_Loss  # noqa: F401
from .core.criterions.w_mse import WeightedMSE  # noqa: F401
from .core.criterions.tversky_loss import FocalTverskyLoss, TverskyLoss  # noqa: F401
from .utils import voxelization  # noqa: F401

__version__ = "0.1.0"
