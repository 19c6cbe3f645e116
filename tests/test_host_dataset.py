"""CPU: the TS40K dataset mirror (file listing, raw __getitem__, transform hook) — no GPU involved."""
import numpy as np
import torch


def test_ts40k_listing_and_raw_items(tmp_path):
    from scenenet_b200.core.datasets.ts40k import TS40K, POWER_LINE_SUPPORT_TOWER
    assert POWER_LINE_SUPPORT_TOWER == 15
    split = tmp_path / "fit"
    split.mkdir()
    rng = np.random.default_rng(0)
    for i, n in enumerate((10, 20, 30)):
        np.save(split / f"sample_{i}.npy", rng.random((n, 4)))
    (split / "readme.txt").write_text("x")
    (split / "sub").mkdir()
    ds = TS40K(str(tmp_path), split="fit")
    assert len(ds) == 3 and sorted(ds.npy_files.tolist()) == ["sample_0.npy", "sample_1.npy", "sample_2.npy"]
    pts, lab = ds[torch.tensor(1)]
    rows = np.load(ds.path_of(1))
    assert pts.shape == (1, rows.shape[0], 3) and lab.shape == (1, rows.shape[0])
    assert np.array_equal(pts[0], rows[:, :3]) and np.array_equal(lab[0], rows[:, 3])
    ds.set_transform(lambda s: (s[0].shape, s[1].shape))
    assert ds[0] == ((np.load(ds.path_of(0)).shape[0], 3), (np.load(ds.path_of(0)).shape[0],))


def test_device_loader_refuses_cpu():
    import pytest
    from scenenet_b200.core.datasets.ts40k import TS40KDeviceLoader
    with pytest.raises((RuntimeError, AssertionError)):
        TS40KDeviceLoader([np.zeros((4, 4))], batch_size=1, device="cpu")
