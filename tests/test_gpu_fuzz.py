"""Randomised shapes (ragged grids, Y % 4 != 0, kernels larger than the grid, occupancy 0 .. 1, both output dtypes):
dense and occupancy-driven forward / tap-gradient kernels and the device-side selection against float64 torch
references (scratch/fuzz_kernels.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 2])
def test_random_shapes(seed):
    env = dict(os.environ, SEED=str(seed), N="60")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scratch", "fuzz_kernels.py")], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "60 cases, 0 problems" in r.stdout, r.stdout[-3000:]
