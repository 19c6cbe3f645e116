"""GPU: the batched device data path (core/datasets/ts40k.py) against the per-sample transform chain
of the reference (Compose([Voxelization([tower], vxg_size), ToTensor(), ToFullDense((True, True))]),
scripts/main.py:135-140) evaluated by the CPU oracle, and against the reference's own golden voxelization of
data-sample/sample_575.npy."""
import os

import numpy as np
import pytest
import torch

from oracle import voxel_oracle as vo

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _cloud(rng, n):
    pts = rng.uniform(0, 30, (n, 3)) + np.array([544850.0, 4634550.0, 160.0])
    lab = rng.choice([1.0, 2.0, 5.0, 15.0, 16.0], n, p=[0.5, 0.3, 0.1, 0.05, 0.05])
    return np.concatenate([pts, lab[:, None]], axis=1)


def _oracle_xy(rows, grid):
    g = vo.raw_grids(rows[:, :3], rows[:, 3], [15], grid)
    return (g["count"] > 0).astype(np.float64), (g["keep"] > 0).astype(np.float64)


@pytest.mark.parametrize("dtype", [torch.float64, torch.uint8])
def test_loader_matches_per_sample_chain(tmp_path, dtype):
    import scenenet_b200 as sb
    rng = np.random.default_rng(5)
    split = tmp_path / "fit"
    split.mkdir()
    clouds = [_cloud(rng, n) for n in (3000, 1, 12345, 800, 64, 20000, 7)]
    for i, c in enumerate(clouds):
        np.save(split / f"sample_{i}.npy", c)
    (split / "notes.txt").write_text("not a sample")
    ds = sb.TS40K(str(tmp_path), split="fit")
    assert len(ds) == len(clouds) and str(ds) == f"TS40K fit Dataset with {len(clouds)} samples"
    raw = ds[0]
    assert raw[0].shape[0] == 1 and raw[0].shape[2] == 3 and raw[1].shape[0] == 1  # (1, N, 3), (1, N) like the reference
    grid = (32, 32, 32)
    loader = sb.TS40KDeviceLoader(ds, batch_size=3, vxg_size=grid, device=DEV, dtype=dtype)
    assert len(loader) == 3
    seen = 0
    for x, y in loader:
        assert x.dtype == dtype and x.dim() == 5 and x.shape[1:] == (1, 32, 32, 32) and x.is_cuda
        for j in range(x.shape[0]):
            rows = np.load(ds.path_of(seen))
            wx, wy = _oracle_xy(rows, grid)
            assert np.array_equal(x[j, 0].cpu().numpy().astype(np.float64), wx), seen
            assert np.array_equal(y[j, 0].cpu().numpy().astype(np.float64), wy), seen
            seen += 1
    assert seen == len(clouds)
    # the model takes the batch as it comes
    model = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, (9, 5, 5)).to(DEV)
    x, y = next(iter(loader))
    pred = model(x)
    assert pred.shape == x.shape and bool(torch.isfinite(pred).all())


def test_loader_golden_sample_575_and_replacement_of_bad_samples():
    import scenenet_b200 as sb
    g = np.load(os.path.join(GOLDEN, "vox_sample_575.npz"))
    s = np.load(os.path.join(GOLDEN, "sample_575.npz"))
    rows = s["npy"]
    empty = np.zeros((0, 4))
    loader = sb.TS40KDeviceLoader([rows, empty, rows], batch_size=3, vxg_size=(64, 64, 64), device=DEV, seed=0)
    x, y = next(iter(loader))
    assert x.shape == (3, 1, 64, 64, 64)
    want_y = np.zeros(64 ** 3)
    want_y[g["ref_frac_idx"]] = 1.0  # the reference's own reg_on_voxel output: 78 tower voxels
    want_y = want_y.reshape(64, 64, 64)
    want_x = np.zeros(64 ** 3)
    want_x[g["restated_count_idx"]] = 1.0
    want_x = want_x.reshape(64, 64, 64)
    for j in range(3):  # the empty sample was replaced by a readable one (all sources but it are sample_575)
        assert np.array_equal(y[j, 0].cpu().numpy(), want_y)
        assert np.array_equal(x[j, 0].cpu().numpy(), want_x) and int(x[j].sum()) == 4247  # SURVEY §8c KAT


def test_in_memory_samples_are_page_locked_in_place():
    """[N, 4] float64 arrays held in memory are registered with the driver once and copied to the device straight from
    the arrays: the batches equal those of the staging path (pin_sources_bytes=0) and of the oracle, twice over (two epochs),
    and a mixed list (one float32 array: not registrable) falls back to staging."""
    import scenenet_b200 as sb
    rng = np.random.default_rng(11)
    clouds = [_cloud(rng, n) for n in (5000, 33, 12000, 900, 64, 7000)]
    grid = (32, 32, 32)
    direct = sb.TS40KDeviceLoader(clouds, batch_size=4, vxg_size=grid, device=DEV)
    staged = sb.TS40KDeviceLoader(clouds, batch_size=4, vxg_size=grid, device=DEV, pin_sources_bytes=0)
    assert not staged._registered
    if not direct._registered:
        pytest.skip("cudaHostRegister is not available through torch.cuda.cudart() on this box")
    assert len(direct._registered) == len(clouds)
    for _ in range(2):
        seen = 0
        for (x, y), (xs, ys) in zip(direct, staged):
            assert torch.equal(x, xs) and torch.equal(y, ys)
            for j in range(x.shape[0]):
                wx, wy = _oracle_xy(clouds[seen], grid)
                assert np.array_equal(x[j, 0].cpu().numpy(), wx) and np.array_equal(y[j, 0].cpu().numpy(), wy)
                seen += 1
        assert seen == len(clouds)
    direct.close()
    assert not direct._registered
    mixed = sb.TS40KDeviceLoader(clouds[:3] + [clouds[3].astype(np.float32)], batch_size=4, vxg_size=grid, device=DEV)
    (x, y), = list(mixed)
    assert torch.equal(x[:3], next(iter(staged))[0][:3])
