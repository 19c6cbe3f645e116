"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "scenenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import scenenet_b200
    from scenenet_b200 import _lib
    names = _header_functions()
    assert len(names) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/scenenet_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib.sn_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in _lib.lib.sn_build_info()


def test_argument_validation_without_gpu():
    """error paths return codes before any CUDA call"""
    from scenenet_b200._lib import lib
    assert lib.sn_scenenet_fwd(None, None, 0, None, 1, 8, 8, 8, 3, 3, 3, None, 0, None) == -1
    assert lib.sn_scenenet_fwd_multi(None, None, 0, None, 3, 1, 8, 8, 8, 3, 3, 3, None, 0, None) == -1
    assert lib.sn_cast_f64_to_f32(None, None, 4, None) == -1
    assert lib.sn_threshold(None, 0, 0.5, 4, None, None) == -1
    assert lib.sn_scenenet_bwd_workspace_bytes(0, 8, 8, 8, 3, 3, 3) == -1
    assert lib.sn_scenenet_bwd_workspace_bytes(32, 64, 64, 64, 9, 5, 5) > 0
    assert lib.sn_grid_prepare(None, 1, 8, None, None, None) == -1
    # ABI v4 layout: 8 counters, one bit per voxel + 4 padding words, the dense-tile list (n / 1024 + 1024 ids)
    assert lib.sn_grid_state_bytes(-1) == -1 and lib.sn_grid_state_bytes(0) == 64 + 16 + 4096
    assert lib.sn_grid_state_bytes(1 << 23) == 64 + (1 << 20) + 16 + 4 * ((1 << 13) + 1024)
    assert lib.sn_confusion_counts(None, 1, None, 3, 8, 0.65, None, None, None) == -1
    assert lib.sn_scenenet_tapgrad(None, None, None, 0, 1, 8, 8, 8, 3, 3, 3, None, None, 0, None) == -1
    assert lib.sn_criterion_workspace_bytes(0) == -1 and lib.sn_criterion_workspace_bytes(1 << 23) > 0
    assert lib.sn_criterion_fwd(None, None, 1, 8, None, None, 10, 1.0, 2.0, 1.0, 4.0, 1e-6, 3, None, None, None, 0, None) == -1
    assert lib.sn_param_penalty(None, None, 0, 5.0, None, None, None) == -1
    assert lib.sn_peer_allreduce_buffer_bytes(8) == 2 * 8 * 256 * 4 and lib.sn_peer_allreduce_buffer_bytes(17) == -1
    assert lib.sn_peer_allreduce(None, 13, 0, 2, None, None, None, 1000, None) == -1
    assert lib.sn_scenenet_param_grads_allreduce(None, None, None, None, None, 1.0, None, 0, 2, None, None, None, 1000, None) == -1
    # the selection rule of the AUTO modes (host-side query): config 2 at 1.6 % -> occupancy-driven forward and backward
    n = int(0.016 * 32 * 64 ** 3)
    assert lib.sn_select_path(0, n, 32, 64, 64, 64, 9, 5, 5) == 2 and lib.sn_select_path(1, n, 32, 64, 64, 64, 9, 5, 5) == 2
    assert lib.sn_select_path(0, 3 * n, 32, 64, 64, 64, 9, 5, 5) == 0           # 4.8 % occupied: per-tile choice (ABI v4)
    assert lib.sn_select_path(0, 20 * n, 32, 64, 64, 64, 9, 5, 5) == 1          # 32 % occupied: dense forward
    assert lib.sn_select_fwd_path_state(n, 0, 32, 64, 64, 64, 9, 5, 5) == 2 and lib.sn_select_fwd_path_state(n, 5000, 32, 64, 64, 64, 9, 5, 5) == 0  # clustered: per tile
    assert lib.sn_select_path(1, 20 * n, 32, 64, 64, 64, 9, 5, 5) == 1          # 32 % occupied: dense tap gradient
    assert lib.sn_select_path(0, n, 8, 128, 128, 128, 9, 9, 9) == 2 and lib.sn_select_path(0, n, 8, 128, 128, 128, 15, 15, 15) == 1


def test_model_desc_layout_matches_header():
    from scenenet_b200._lib import ModelDesc, SN_MAX_GENEOS
    assert ctypes.sizeof(ModelDesc) == 4 * (5 + 4 * SN_MAX_GENEOS + 1)


def test_no_cpu_fallback_on_cpu_tensors():
    import pytest
    import torch
    import scenenet_b200 as sb
    torch.manual_seed(0)
    m = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 8, 8, 8, dtype=torch.float64))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "scene-net_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"


def test_state_dict_keys_and_rng_stream_match_reference_layout():
    import torch
    import scenenet_b200 as sb
    torch.manual_seed(0)
    m = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5))
    keys = list(m.state_dict().keys())
    assert keys == ['geneos.cy_0.geneo_params.radius', 'geneos.cy_0.geneo_params.sigma',
                    'geneos.cone_0.geneo_params.apex', 'geneos.cone_0.geneo_params.cone_inc',
                    'geneos.cone_0.geneo_params.cone_radius', 'geneos.cone_0.geneo_params.radius',
                    'geneos.cone_0.geneo_params.sigma', 'geneos.neg_0.geneo_params.neg_factor',
                    'geneos.neg_0.geneo_params.radius', 'geneos.neg_0.geneo_params.sigma',
                    'lambdas_dict.lambda_cone_0', 'lambdas_dict.lambda_cy_0', 'lambdas_dict.lambda_neg_0']
    assert m.last_lambda == 'lambda_cy_0'  # what the reference picks under seed 0 (tests/golden/meta.json)
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert frozen == ['geneos.cone_0.geneo_params.apex', 'lambdas_dict.lambda_cy_0']
    assert m.get_num_total_params() == 11
    assert abs(float(sum(m.lambdas_dict.values())) - 1.0) < 1e-6


import pytest


@pytest.mark.reference
def test_constructor_matches_live_reference():
    import torch
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_shim
    ref_shim.install()
    from core.models.SCENE_Net import SceneNet as RefSceneNet
    import scenenet_b200 as sb
    for seed, gn, ks in [(0, {'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5)), (3, {'cy': 2, 'cone': 1, 'neg': 2}, (9, 7, 7))]:
        torch.manual_seed(seed)
        r = RefSceneNet(dict(gn), ks)
        torch.manual_seed(seed)
        m = sb.SceneNet(dict(gn), ks)
        assert r.last_lambda == m.last_lambda
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        for k in rs:
            assert float(rs[k]) == float(ms[k]), k
        assert [p.requires_grad for p in r.parameters()] == [p.requires_grad for p in m.parameters()]
