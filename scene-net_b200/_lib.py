"""ctypes binding of the C-ABI CUDA library (include/scenenet_b200.h).

There is deliberately NO fallback: if `libscenenet_b200.so` is missing or a symbol is
absent the import of the ops fails loudly.  Build it with `python scene-net_b200/build.py`
(or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCENENET_B200_LIB", os.path.join(PKG, "libscenenet_b200.so"))  # override: experiments only

SN_F32, SN_F64, SN_U8, SN_I32, SN_I64, SN_BITS = 0, 1, 2, 3, 4, 5
SN_PATH_AUTO, SN_PATH_DENSE, SN_PATH_SPARSE = 0, 1, 2
SN_TAPGRAD_AUTO, SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE = 0, 1, 2
SN_MAX_GENEOS = 16
SN_MAX_PARAM_PTRS = 96
SN_MAX_TAPS = 4096
SN_CRIT_MAX_BINS = 12
SN_CRIT_COEF = SN_CRIT_MAX_BINS + 4
ABI_VERSION = 5

KIND = {
    "cylinder_kernel": 0, "cylinderv2": 1, "cone_kernel": 2, "arrow": 3, "neg_sphere_kernel": 4, "negSpherev2": 5,
}


class ModelDesc(C.Structure):
    """mirror of `sn_model_desc`"""
    _fields_ = [
        ("n_geneos", C.c_int32), ("kz", C.c_int32), ("kx", C.c_int32), ("ky", C.c_int32),
        ("n_param_ptrs", C.c_int32),
        ("kind", C.c_int32 * SN_MAX_GENEOS),
        ("param_index", C.c_int32 * SN_MAX_GENEOS),
        ("lambda_index", C.c_int32 * SN_MAX_GENEOS),
        ("lambda_sum_order", C.c_int32 * SN_MAX_GENEOS),
        ("last_lambda", C.c_int32),
    ]


_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_descp = C.POINTER(ModelDesc)
_pp = C.POINTER(C.c_void_p)
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); must list every function include/scenenet_b200.h declares
SIGNATURES = {
    "sn_abi_version": (_i, []),
    "sn_build_info": (C.c_char_p, []),
    "sn_launch_count": (_i64, []),
    "sn_geneo_synth_fwd": (_i, [_descp, _pp, _vp, _vp, _vp, _vp, _i, _vp]),
    "sn_geneo_synth_bwd": (_i, [_descp, _pp, _vp, _vp, _vp]),
    "sn_scenenet_param_grads": (_i, [_descp, _pp, _vp, _vp, _vp, _d, _vp, _vp]),
    "sn_scenenet_fwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "sn_scenenet_fwd_multi": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "sn_scenenet_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "sn_scenenet_bwd": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "sn_select_path": (_i, [_i, _i64, _i, _i, _i, _i, _i, _i, _i]),
    "sn_select_fwd_path": (_i, [_i64, _i, _i, _i, _i, _i, _i, _i]),
    "sn_select_fwd_path_state": (_i, [_i64, _i64, _i, _i, _i, _i, _i, _i, _i]),
    "sn_scenenet_g0": (_i, [_vp, _i, _vp, _i, _i64, _vp, _vp]),
    "sn_scenenet_tapgrad_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "sn_scenenet_tapgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "sn_grid_state_bytes": (_i64, [_i64]),
    "sn_grid_prepare": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "sn_criterion_workspace_bytes": (_i64, [_i64]),
    "sn_criterion_fwd": (_i, [_vp, _vp, _i, _i64, _fp, _fp, _i, C.c_float, _d, _d, _d, _d, _i, _vp, _vp, _vp, _i64, _vp]),
    "sn_criterion_bwd": (_i, [_vp, _vp, _i, _i64, _fp, _fp, _i, _vp, _vp, _vp, _i, _vp]),
    "sn_param_penalty": (_i, [_pp, C.POINTER(C.c_int32), _i, C.c_float, _vp, _vp, _vp]),
    "sn_cast_f64_to_f32": (_i, [_vp, _vp, _i64, _vp]),
    "sn_cast_u8_to_f32": (_i, [_vp, _vp, _i64, _vp]),
    "sn_threshold": (_i, [_vp, _i, _d, _i64, _vp, _vp]),
    "sn_vxg_to_xyz": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sn_confusion_counts": (_i, [_vp, _i, _vp, _i, _i64, _d, _vp, _vp, _vp]),
    "sn_vox_minmax": (_i, [_vp, _i, _vp, _i, _i64, _vp, _vp]),
    "sn_vox_edges": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sn_vox_bin": (_i, [_vp, _i, _vp, _i, _vp, _i, _i64, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sn_vox_voxelize": (_i, [_vp, _i, _vp, _i, _vp, _i, _i64, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sn_vox_finalize_workspace_bytes": (_i64, [_i, _i]),
    "sn_vox_finalize": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sn_peer_allreduce_buffer_bytes": (_i64, [_i]),
    "sn_peer_allreduce": (_i, [_vp, _i, _i, _i, C.POINTER(C.c_uint64), _vp, _vp, _i64, _vp]),
    "sn_scenenet_param_grads_allreduce": (_i, [_descp, _pp, _vp, _vp, _vp, _d, _vp, _i, _i, C.POINTER(C.c_uint64), _vp, _vp, _i64, _vp]),
    "sn_fp32_peak_probe": (_i, [_vp, _i, C.POINTER(C.c_double), _vp]),
}

_ERRORS = {-1: "SN_ERR_BAD_ARG", -2: "SN_ERR_UNSUPPORTED", -3: "SN_ERR_ALIGN", -4: "SN_ERR_WORKSPACE"}


class SceneNetB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the scenenet_b200 CUDA library is not built and there is no CPU fallback. "
            f"Run `python {os.path.join(PKG, 'build.py')}` (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    v = lib.sn_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {v}, python binding {ABI_VERSION}; rebuild the library")
    return lib


lib = _load()


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc <= -1000:
        raise SceneNetB200Error(f"{what}: CUDA error {-(rc + 1000)}")
    raise SceneNetB200Error(f"{what}: {_ERRORS.get(rc, rc)}")


def launch_count() -> int:
    return int(lib.sn_launch_count())
