"""GPU parity tests of the voxelization path through the C ABI: bit-exact voxel indices,
counts, keep counts, max labels, densities and fractions against the numpy oracle, the
reference's own reg_on_voxel output (golden) and domain properties at BASELINE sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import voxel_oracle as vo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _vox():
    from scenenet_b200.utils import voxelization as Vox
    return Vox


def _dense(idx, val, n):
    a = np.zeros(n, dtype=val.dtype)
    a[idx] = val
    return a


@pytest.fixture(scope="module")
def s575(golden_dir):
    npy = np.load(os.path.join(golden_dir, "sample_575.npz"))["npy"]
    gold = np.load(os.path.join(golden_dir, "vox_sample_575.npz"))
    return npy, gold


def test_sample_575_bit_exact(s575):
    from scenenet_b200 import voxel_ops
    npy, gold = s575
    full = torch.from_numpy(npy).to(DEV)          # [N,4] rows x,y,z,label as stored by TS40K
    out = voxel_ops.voxelize_clouds(full[:, :3], None, (64, 64, 64), full[:, 3], [15],
                                    want=("density", "frac", "max_label", "occ", "occ_keep", "keep_count"),
                                    occ_dtype=torch.float64, return_lin=True)
    V = 64 ** 3
    assert np.array_equal(out["lin"].cpu().numpy(), gold["restated_lin"])
    assert np.array_equal(out["edges"].cpu().numpy().reshape(3, 65), gold["edges"])
    assert np.array_equal(out["count"].cpu().numpy().reshape(-1), _dense(gold["restated_count_idx"], gold["restated_count_val"], V))
    assert np.array_equal(out["frac"].cpu().numpy().reshape(-1), _dense(gold["ref_frac_idx"], gold["ref_frac_val"], V))  # reference's own output
    assert np.array_equal(out["max_label"].cpu().numpy().reshape(-1), _dense(gold["restated_maxlab_idx"], gold["restated_maxlab_val"], V))
    assert np.array_equal(out["density"].cpu().numpy().reshape(-1), _dense(gold["restated_density_idx"], gold["restated_density_val"], V))
    assert np.array_equal(out["occ"].cpu().numpy().reshape(-1) > 0, _dense(gold["restated_count_idx"], gold["restated_count_val"], V) > 0)
    assert int(out["occ_keep"].sum()) == 78 and int(out["count"].sum()) == 58243 and int(out["count"].max()) == 81
    assert float(out["frac"].sum()) == 61.60299032173086


def test_reference_api_functions(s575):
    Vox = _vox()
    npy, gold = s575
    pts, labels = npy[:, 0:-1], npy[:, -1]          # non-contiguous views, as ts40k.py:207 hands them over
    f128 = Vox.reg_on_voxel(pts, labels, [15], voxelgrid_dims=(128, 128, 128))
    assert np.array_equal(f128.reshape(-1), _dense(gold["ref_frac128_idx"], gold["ref_frac128_val"], 128 ** 3))
    d = Vox.hist_on_voxel(pts, voxelgrid_dims=(64, 64, 64))
    assert d.shape == (64, 64, 64) and d.dtype == np.float64
    assert np.array_equal(d, vo.hist_on_voxel(pts, (64, 64, 64)))
    c = Vox.classes_on_voxel(pts, labels, (32, 32, 32))
    assert np.array_equal(c, vo.classes_on_voxel(pts, labels, (32, 32, 32)))
    f = Vox.reg_on_voxel(pts, labels, 15, voxelgrid_dims=(32, 48, 16))   # scalar keep label, anisotropic grid
    assert f.shape == (16, 32, 48)
    assert np.array_equal(f, vo.reg_on_voxel(pts, labels, 15, (32, 48, 16)))


def test_transform_chain(s575):
    import scenenet_b200 as sb
    npy, _ = s575
    sample = (npy[:, 0:-1], npy[:, -1])
    d, f = sb.Voxelization([15], vxg_size=(64, 64, 64))(sample)
    do, fo = vo.voxelization_transform(sample[0], sample[1], [15], vxg_size=(64, 64, 64))
    assert d.shape == (1, 64, 64, 64) and np.array_equal(d, do) and np.array_equal(f, fo)
    x, y = sb.ToFullDense(apply=(True, True))(sb.ToTensor()((d, f)))
    assert x.dtype == torch.float64 and int(x.sum()) == 4247 and int(y.sum()) == 78


def test_size_mode(s575):
    Vox = _vox()
    npy, _ = s575
    pts, labels = npy[:, :3], npy[:, 3]
    for vd in [(1.0, 1.0, 1.0), (2.0, 2.0, 0.5)]:
        f = Vox.reg_on_voxel(pts, labels, [15], voxel_dims=vd)
        assert np.array_equal(f, vo.reg_on_voxel(pts, labels, [15], voxel_dims=vd))
        d = Vox.hist_on_voxel(pts, voxel_dims=vd)
        assert np.array_equal(d, vo.hist_on_voxel(pts, voxel_dims=vd))


def _synthetic_cloud(n, seed, utm=True):
    """ground sheet + vegetation blobs + tower column, UTM offset."""
    rng = np.random.default_rng(seed)
    n_g, n_v = int(0.70 * n), int(0.25 * n)
    n_t = n - n_g - n_v
    ground = np.column_stack([rng.uniform(0, 30, n_g), rng.uniform(0, 30, n_g), rng.normal(0, 0.3, n_g)])
    centres = rng.uniform(0, 30, (20, 3)) * np.array([1, 1, 0.3])
    veg = centres[rng.integers(0, 20, n_v)] + rng.normal(0, 2.0, (n_v, 3))
    tower = np.column_stack([rng.normal(15, 0.5, n_t), rng.normal(15, 0.5, n_t), rng.uniform(0, 40, n_t)])
    pts = np.concatenate([ground, veg, tower])
    labels = np.concatenate([rng.choice([1, 2, 3, 4, 5, 7, 10, 11, 12, 16, 19], n_g + n_v), np.full(n_t, 15)]).astype(np.float64)
    if utm:
        pts = pts + np.array([544850.0, 4634550.0, 160.0])
    perm = rng.permutation(n)
    return pts[perm], labels[perm]


@pytest.mark.parametrize("n,grid", [(100_000, (64, 64, 64)), (300_000, (128, 128, 128)), (1_000_000, (64, 64, 64))])
def test_synthetic_clouds_bit_exact(n, grid):
    from scenenet_b200 import voxel_ops
    pts, labels = _synthetic_cloud(n, seed=n % 97)
    g = vo.raw_grids(pts, labels, [15], grid)
    out = voxel_ops.voxelize_clouds(torch.from_numpy(pts).to(DEV), None, grid, torch.from_numpy(labels).to(DEV), [15],
                                    want=("frac", "max_label", "keep_count", "density"), return_lin=True)
    assert np.array_equal(out["lin"].cpu().numpy(), g["lin"])
    assert np.array_equal(out["count"][0].cpu().numpy(), g["count"])
    assert np.array_equal(out["keep_count"][0].cpu().numpy(), g["keep"])
    assert np.array_equal(out["max_label"][0].cpu().numpy(), g["maxlab"])
    assert np.array_equal(out["density"][0].cpu().numpy(), vo.normalize_minmax(g["count"].astype(np.float64)))


def test_edge_cases():
    from scenenet_b200 import voxel_ops
    # points exactly on edges, duplicates, min/max corners
    pts = np.array([[0.0, 0.0, 0.0], [4.0, 4.0, 4.0], [1.0, 2.0, 3.0], [1.0 + 1e-12, 2.0, 3.0 - 1e-12], [1.0, 2.0, 3.0],
                    [2.0, 2.0, 2.0], [4.0, 0.0, 2.0]])
    labels = np.array([1.0, 15.0, 15.0, 2.0, -3.0, 7.0, 15.0])
    for grid in [(4, 4, 4), (3, 5, 2), (1, 1, 1)]:
        g = vo.raw_grids(pts, labels, [15, 7], grid)
        out = voxel_ops.voxelize_clouds(torch.from_numpy(pts).to(DEV), None, grid, torch.from_numpy(labels).to(DEV), [15, 7],
                                        want=("frac", "max_label", "keep_count"), return_lin=True)
        assert np.array_equal(out["lin"].cpu().numpy(), g["lin"]), grid
        assert np.array_equal(out["count"][0].cpu().numpy(), g["count"])
        assert np.array_equal(out["keep_count"][0].cpu().numpy(), g["keep"])
        assert np.array_equal(out["max_label"][0].cpu().numpy(), g["maxlab"])
    # a single point / all points identical (zero range): everything lands in voxel 0
    one = np.array([[5.0, 6.0, 7.0]] * 3)
    out = voxel_ops.voxelize_clouds(torch.from_numpy(one).to(DEV), None, (4, 4, 4), want=(), return_lin=True)
    g = vo.raw_grids(one, voxelgrid_dims=(4, 4, 4))
    assert np.array_equal(out["lin"].cpu().numpy(), g["lin"]) and int(out["count"].sum()) == 3
    # negative labels only: max label stays negative, empty voxels are 0
    neg = vo.raw_grids(pts, -np.abs(labels) - 1, None, (2, 2, 2))
    outn = voxel_ops.voxelize_clouds(torch.from_numpy(pts).to(DEV), None, (2, 2, 2), torch.from_numpy(-np.abs(labels) - 1).to(DEV),
                                     want=("max_label",))
    assert np.array_equal(outn["max_label"][0].cpu().numpy(), neg["maxlab"])


def test_batched_ragged_clouds_and_row_layout():
    """several clouds of different sizes in ONE set of launches; [N,4] rows with the label in the row"""
    from scenenet_b200 import voxel_ops
    sizes = [50_000, 1, 120_000, 777]
    clouds = [_synthetic_cloud(max(n, 4), seed=i) for i, n in enumerate(sizes)]
    clouds = [(p[:n], l[:n]) for (p, l), n in zip(clouds, sizes)]
    rows = np.concatenate([np.column_stack([p, l]) for p, l in clouds])
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    full = torch.from_numpy(rows).to(DEV)
    out = voxel_ops.voxelize_clouds(full[:, :3], torch.from_numpy(offs).to(DEV), (64, 64, 64), full[:, 3], [15],
                                    want=("frac", "occ", "occ_keep", "keep_count", "density"))
    for c, (p, l) in enumerate(clouds):
        g = vo.raw_grids(p, l, [15], (64, 64, 64))
        assert np.array_equal(out["count"][c].cpu().numpy(), g["count"]), c
        assert np.array_equal(out["keep_count"][c].cpu().numpy(), g["keep"]), c
        assert np.array_equal(out["frac"][c].cpu().numpy(), vo.reg_on_voxel(p, l, [15], (64, 64, 64))), c
        assert np.array_equal(out["density"][c].cpu().numpy(), vo.hist_on_voxel(p, (64, 64, 64))), c
        assert np.array_equal(out["occ"][c].cpu().numpy() > 0, g["count"] > 0)


def test_properties_at_10m_points():
    """BASELINE config 3 top size: size-independent properties (sum of counts, keep <= count, idempotence)."""
    from scenenet_b200 import voxel_ops
    n = 10_000_000
    g = torch.Generator(device=DEV).manual_seed(3)
    pts = torch.rand((n, 3), generator=g, device=DEV, dtype=torch.float64) * torch.tensor([30.0, 30.0, 40.0], device=DEV, dtype=torch.float64) \
        + torch.tensor([544850.0, 4634550.0, 160.0], device=DEV, dtype=torch.float64)
    labels = torch.randint(1, 20, (n,), generator=g, device=DEV).to(torch.float64)
    for grid in [(64, 64, 64), (128, 128, 128)]:
        o1 = voxel_ops.voxelize_clouds(pts, None, grid, labels, [15], want=("keep_count", "max_label"), return_lin=True)
        assert int(o1["count"].sum()) == n
        assert int(o1["keep_count"].sum()) == int((labels == 15).sum())
        assert bool((o1["keep_count"] <= o1["count"]).all())
        lin = o1["lin"].to(torch.int64)
        assert int(lin.min()) >= 0 and int(lin.max()) < grid[0] * grid[1] * grid[2]
        assert torch.equal(torch.bincount(lin, minlength=grid[0] * grid[1] * grid[2]).view_as(o1["count"][0]).to(torch.int32), o1["count"][0])
        o2 = voxel_ops.voxelize_clouds(pts, None, grid, labels, [15], want=("keep_count", "max_label"))
        assert torch.equal(o1["count"], o2["count"]) and torch.equal(o1["max_label"], o2["max_label"])
        assert float(o1["max_label"].max()) == 19.0


def test_vxg_to_xyz_vs_reference_outputs_and_oracle():
    """vxg_to_xyz (utils/voxelization.py:328-360): bit-exact against the reference's own outputs on the committed fixtures and
    against the oracle on a 64^3 prediction-shaped grid with UTM origin (every voxel, C order, float64)."""
    import scenenet_b200 as sb
    from oracle import voxel_oracle as vo
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vxg_to_xyz.npz"))
    for name in sorted({k.split(".")[0] for k in gold.files}):
        origin = gold[f"{name}.origin"] if gold[f"{name}.origin"].size else None
        size = gold[f"{name}.size"] if gold[f"{name}.size"].size else None
        vxg = gold[f"{name}.vxg"]
        for arg in (vxg, torch.from_numpy(vxg), torch.from_numpy(vxg).cuda()):
            out = sb.voxelization.vxg_to_xyz(arg, origin, size)
            assert isinstance(out, np.ndarray) and out.dtype == np.float64
            assert out.shape == gold[f"{name}.out"].shape and np.array_equal(out, gold[f"{name}.out"]), name
    rng = np.random.default_rng(3)
    grid = (rng.random((64, 64, 64)) < 0.02).astype(np.float64)
    origin, size = np.array([544850.123, 4634550.456, 160.789]), np.array([0.46875, 0.46875, 0.7])
    out = sb.voxelization.vxg_to_xyz(grid, origin, size)
    assert np.array_equal(out, vo.vxg_to_xyz(grid, origin, size))
    # the rows a caller keeps: label == 1
    assert int((out[:, 3] == 1.0).sum()) == int(grid.sum())
    with pytest.raises(ValueError):
        sb.voxelization.vxg_to_xyz(np.zeros((2, 3, 4, 5)))
    assert sb.voxelization.vxg_to_xyz(np.zeros((0, 3, 4))).shape == (0, 4)
