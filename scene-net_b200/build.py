"""Builds the in-tree C-ABI CUDA library `libscenenet_b200.so` for sm_100a with nvcc.

No torch involved: the library's ABI is plain C (include/scenenet_b200.h) and it links only
the static CUDA runtime, so it loads on a box without a GPU (symbol checks) and travels to
the GPU box inside the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
BUILD = os.path.join(ROOT, "build", "scenenet_b200")
LIB = os.path.join(PKG, "libscenenet_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extra flags: the synthesis follows the reference's float32 op order, no contraction
EXTRA = {"synth.cu": ["-fmad=false"]}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the scenenet_b200 CUDA library cannot be built")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(src: str) -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/scenenet_b200.h"]:
        p = os.path.join(CSRC, f)
        if f.endswith((".cuh", ".h")) or f == src:
            with open(p, "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(ARCH + COMMON + EXTRA.get(src, [])).encode())
    return h.hexdigest()


def _compile(src: str, force: bool) -> tuple[str, str]:
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    stamp = obj + ".sha"
    dg = _digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dg:
        return obj, ""
    cmd = [_nvcc(), *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dg)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in res]
    log = "\n".join(l for _, l in res if l)
    if log:
        with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        # --no-undefined: a declaration that no longer matches its definition must fail HERE, not as an unresolved symbol
        # when the library is loaded on the GPU box
        cmd = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
