"""TEST INFRASTRUCTURE ONLY — tests/golden/ref_vxg_to_xyz.npz by RUNNING THE REAL REFERENCE's vxg_to_xyz
(utils/voxelization.py:328-360) on small grids.  Run in the build container:  python -m oracle.make_golden_vxg"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    ref_shim.install()
    from utils import voxelization as Vox
    rng = np.random.default_rng(7)
    cases = {
        "f64_default": (torch.from_numpy((rng.random((4, 3, 5)) < 0.3).astype(np.float64)), None, None),
        "f32_origin_size": (torch.from_numpy(rng.random((3, 4, 2)).astype(np.float32)), np.array([544850.25, 4634550.5, 160.0]),
                            np.array([0.5, 0.25, 1.5])),
        "u8_int_origin": (torch.from_numpy((rng.random((2, 2, 6)) < 0.5).astype(np.uint8)), np.array([3, -2, 7]), np.array([2, 2, 2])),
    }
    out = {}
    for name, (vxg, origin, size) in cases.items():
        res = np.asarray(Vox.vxg_to_xyz(vxg, origin, size))
        out[f"{name}.vxg"] = vxg.numpy()
        out[f"{name}.origin"] = np.zeros(0) if origin is None else origin.astype(np.float64)
        out[f"{name}.size"] = np.zeros(0) if size is None else size.astype(np.float64)
        out[f"{name}.out"] = res.astype(np.float64)
        print(name, res.shape, res.dtype)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_vxg_to_xyz.npz"), **out)


if __name__ == "__main__":
    main()
