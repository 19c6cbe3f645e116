"""Pins oracle/voxel_oracle.py: against the reference's own reg_on_voxel output (golden),
sklearn's MinMaxScaler, np.linspace / np.searchsorted, and the SURVEY §8c KAT.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import voxel_oracle as vo


@pytest.fixture(scope="module")
def s575(golden_dir):
    npy = np.load(os.path.join(golden_dir, "sample_575.npz"))["npy"]
    gold = np.load(os.path.join(golden_dir, "vox_sample_575.npz"))
    meta = json.load(open(os.path.join(golden_dir, "meta.json")))
    return npy, gold, meta


def _dense(idx, val, n):
    a = np.zeros(n, dtype=val.dtype)
    a[idx] = val
    return a


def test_reg_on_voxel_matches_reference_output(s575):
    npy, gold, meta = s575
    pts, labels = npy[:, :3], npy[:, 3]
    f = vo.reg_on_voxel(pts, labels, [15], (64, 64, 64))
    ref = _dense(gold["ref_frac_idx"], gold["ref_frac_val"], 64 ** 3).reshape(64, 64, 64)
    assert np.array_equal(f, ref)  # bit-exact (float64 true division of the same integers)
    f128 = vo.reg_on_voxel(pts, labels, [15], (128, 128, 128))
    ref128 = _dense(gold["ref_frac128_idx"], gold["ref_frac128_val"], 128 ** 3).reshape(128, 128, 128)
    assert np.array_equal(f128, ref128)


def test_survey_kat_numbers(s575):
    npy, gold, meta = s575
    g = vo.raw_grids(npy[:, :3], npy[:, 3], [15], (64, 64, 64))
    assert len(npy) == 58243
    assert int((g["count"] > 0).sum()) == 4247
    assert int(g["count"].max()) == 81
    assert int(g["count"].sum()) == 58243
    assert int(g["lin"].sum()) == 2389522047
    assert int((g["keep"] > 0).sum()) == 78
    f = vo.reg_on_voxel(npy[:, :3], npy[:, 3], [15], (64, 64, 64))
    assert f.sum() == 61.60299032173086
    assert np.allclose(g["vg"]["xyzmin"], [544834.005, 4634520.86, 159.49])
    assert np.allclose(g["vg"]["xyzmax"], [544898.085, 4634584.94, 223.57])
    assert np.array_equal(g["lin"], gold["restated_lin"])
    assert np.array_equal(g["count"].reshape(-1), _dense(gold["restated_count_idx"], gold["restated_count_val"], 64 ** 3))


def test_normalize_equals_sklearn():
    from sklearn.preprocessing import MinMaxScaler
    rng = np.random.default_rng(0)
    for shp in [(8, 8, 8), (16, 4, 32), (5, 7, 3)]:
        d = rng.integers(0, 50, size=shp).astype(np.float64)
        d[:, :, 0] = 3.0  # a zero-range column
        ref = MinMaxScaler().fit_transform(d.reshape(-1, shp[-1])).reshape(shp)
        assert np.array_equal(vo.normalize_minmax(d), ref)


def test_linspace_restatement_is_numpy_linspace(s575):
    npy, gold, _ = s575
    for n in (64, 128, 7, 1):
        for i in range(3):
            lo, hi = gold["xyzmin"][i], gold["xyzmax"][i]
            assert np.array_equal(vo.linspace_edges(lo, hi, n), np.linspace(lo, hi, n + 1))
    assert np.array_equal(vo.linspace_edges(2.0, 2.0, 4), np.linspace(2.0, 2.0, 5))


def test_index_rule_edges():
    """point on an interior edge -> lower voxel; min -> 0; max -> n-1."""
    pts = np.array([[0.0, 0.0, 0.0], [4.0, 4.0, 4.0], [1.0, 2.0, 3.0], [1.0 + 1e-12, 2.0, 3.0 - 1e-12]])
    vg = vo.voxelgrid_compute(pts, n_x=4, n_y=4, n_z=4)
    assert vg["voxel_x"].tolist() == [0, 3, 0, 1]
    assert vg["voxel_y"].tolist() == [0, 3, 1, 1]
    assert vg["voxel_z"].tolist() == [0, 3, 2, 2]


def test_size_mode_shapes():
    rng = np.random.default_rng(1)
    pts = rng.uniform(0, 10, size=(1000, 3)) * np.array([1.0, 0.5, 0.25])
    vg = vo.voxelgrid_compute(pts, size_x=0.5, size_y=0.5, size_z=0.2)
    n = vg["x_y_z"]
    assert all(v >= 1 for v in n)
    assert vg["voxel_x"].max() < n[0] and vg["voxel_y"].max() < n[1] and vg["voxel_z"].max() < n[2]
    g = vo.raw_grids(pts, voxel_dims=(0.5, 0.5, 0.2))
    assert g["count"].sum() == 1000 and g["count"].shape == (n[2], n[0], n[1])


def test_classes_and_transform(s575):
    npy, gold, _ = s575
    pts, labels = npy[:, :3], npy[:, 3]
    c = vo.classes_on_voxel(pts, labels, (64, 64, 64))
    assert np.array_equal(c.reshape(-1), _dense(gold["restated_maxlab_idx"], gold["restated_maxlab_val"], 64 ** 3))
    d, f = vo.voxelization_transform(pts, labels, [15], vxg_size=(64, 64, 64))
    assert d.shape == (1, 64, 64, 64) and f.shape == (1, 64, 64, 64)
    assert np.array_equal(vo.densify(d).reshape(-1) > 0, _dense(gold["restated_count_idx"], gold["restated_count_val"], 64 ** 3) > 0)
    assert np.array_equal(vo.prob_to_label(f, 0.5), (f >= 0.5).astype(f.dtype))


@pytest.mark.reference
def test_live_reference_reg_on_voxel_other_samples():
    """Build container only: the reference's reg_on_voxel on other data-sample clouds."""
    from oracle import ref_shim
    ref_shim.install()
    from utils import voxelization as Vox
    import glob
    files = sorted(glob.glob(os.path.join(ref_shim.REF_ROOT, "data-sample", "sample_*.npy")))[:3]
    for fpath in files:
        npy = np.load(fpath)
        pts, labels = npy[:, :3], npy[:, 3]
        ref = Vox.reg_on_voxel(pts, labels, [15], voxelgrid_dims=(32, 32, 32))
        assert np.array_equal(vo.reg_on_voxel(pts, labels, [15], (32, 32, 32)), ref)


def test_vxg_to_xyz_restatement_equals_reference_outputs():
    """oracle.voxel_oracle.vxg_to_xyz against outputs of the reference's own function (oracle/make_golden_vxg.py)."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vxg_to_xyz.npz"))
    names = sorted({k.split(".")[0] for k in gold.files})
    assert names == ["f32_origin_size", "f64_default", "u8_int_origin"]
    for name in names:
        origin = gold[f"{name}.origin"] if gold[f"{name}.origin"].size else None
        size = gold[f"{name}.size"] if gold[f"{name}.size"].size else None
        out = vo.vxg_to_xyz(gold[f"{name}.vxg"], origin, size)
        assert out.shape == gold[f"{name}.out"].shape and np.array_equal(out, gold[f"{name}.out"]), name
