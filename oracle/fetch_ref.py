"""TEST INFRASTRUCTURE ONLY — puts an UNMODIFIED copy of the reference files this path touches under
`oracle/_ref/` so that the real reference travels to the GPU box.

`/root/reference` exists only in the build container.  `oracle/_ref/` is git-ignored (never part of the
history — the reference's sources are not copied into the repository) but NOT gpurun-ignored, so `gpurun` and
the driver's round-end snapshot ship it like the built `.so`.  `__graft_entry__.build()` runs this whenever
`/root/reference` is present; `oracle/ref_shim.py` falls back to `oracle/_ref` when `/root/reference` is
absent.  With it the `-m gpu` tests run the reference's own `GENEO_Tversky_Loss`, `LitSceneNet.training_step`
and Lightning checkpoints on top of the CUDA modules, and `bench.py --impl reference` times the real
reference (`cpu_baseline.kind: "reference"`).

What is copied (byte for byte, checked by sha256 in MANIFEST.json):
  core/**.py, core/criterions/hist_estimation.pickle      the model, GENEO kernels, criterions, Lightning wrappers
  utils/*.py, scripts/*.py                                voxelization / pcd_processing / scripts_utils / constants
  data-sample/sample_575.npy (+ 577, 593)                 the config-1 cloud and two more TS40K samples
  experiments/.../checkpoints/{FBetaScore,JaccardIndex,last,train_loss}.ckpt   Lightning checkpoints (16 KB each)

Usage:  python -m oracle.fetch_ref [--src /root/reference]
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "oracle", "_ref")
DEFAULT_SRC = "/root/reference"

CKPT_DIR = "experiments/scenenet_ts40k/wandb/run-20230217_161733-bwsbqxgs/files/checkpoints"
PATTERNS = [
    "core/**/*.py", "core/criterions/hist_estimation.pickle",
    "utils/*.py", "scripts/*.py", "scripts/*.yml",
    "data-sample/sample_575.npy", "data-sample/sample_577.npy", "data-sample/sample_593.npy",
    f"{CKPT_DIR}/FBetaScore.ckpt", f"{CKPT_DIR}/JaccardIndex.ckpt", f"{CKPT_DIR}/last.ckpt", f"{CKPT_DIR}/train_loss.ckpt",
    "experiments/scenenet_ts40k/*.yml", "requirements.txt",
]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def fetch(src: str = DEFAULT_SRC, dst: str = DST, verbose: bool = False) -> dict:
    if not os.path.isdir(os.path.join(src, "core", "models")):
        raise RuntimeError(f"no reference tree at {src}")
    manifest = {}
    for pat in PATTERNS:
        for p in sorted(glob.glob(os.path.join(src, pat), recursive=True)):
            rel = os.path.relpath(p, src)
            out = os.path.join(dst, rel)
            digest = _sha(p)
            manifest[rel] = digest
            if os.path.exists(out) and _sha(out) == digest:
                continue
            os.makedirs(os.path.dirname(out), exist_ok=True)
            if os.path.exists(out):
                os.chmod(out, 0o644)
            shutil.copyfile(p, out)
            os.chmod(out, 0o644)
            if verbose:
                print("copied", rel)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dst: str = DST) -> bool:
    """True when oracle/_ref holds every file of its manifest with the recorded digest."""
    mf = os.path.join(dst, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    with open(mf) as f:
        files = json.load(f)["files"]
    return all(os.path.exists(os.path.join(dst, rel)) and _sha(os.path.join(dst, rel)) == d for rel, d in files.items())


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else DEFAULT_SRC
    m = fetch(src, verbose=True)
    size = sum(os.path.getsize(os.path.join(DST, r)) for r in m)
    print(f"oracle/_ref: {len(m)} files, {size / 1e6:.1f} MB, verified={verify()}")
