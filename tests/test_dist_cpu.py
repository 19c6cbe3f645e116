"""CPU (gloo, world size 2): the host-side logic of the multi-GPU path — batch sharding and the
single all-reduce of the parameter-gradient payload with DDP mean semantics."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from scenenet_b200 import dist as sdist
    r, w, dev = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    lo, hi = sdist.shard_range(7, r, w)
    params = [torch.nn.Parameter(torch.zeros(())) for _ in range(5)]
    for i, p in enumerate(params[:4]):          # the last one has no gradient (frozen / unused)
        p.grad = torch.tensor(float((rank + 1) * (i + 1)))
    sdist.allreduce_mean_grads(params)
    mean = [float(p.grad) for p in params[:4]]
    # pre-scaled variant: every rank already multiplied by 1/world, the collective only sums
    for i, p in enumerate(params[:4]):
        p.grad = torch.tensor(float((rank + 1) * (i + 1)) / world)
    sdist.allreduce_mean_grads(params, already_scaled=True)
    pre = [float(p.grad) for p in params[:4]]
    out.put((rank, lo, hi, mean, pre, params[4].grad is None))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, m0, p0, n0), (r1, lo1, hi1, m1, p1, n1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)                     # contiguous, balanced, first ranks get the remainder
    expect = [1.5 * (i + 1) for i in range(4)]                      # mean over ranks of (rank+1)*(i+1)
    assert m0 == m1 == expect and p0 == p1 == expect
    assert n0 and n1


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from scenenet_b200.dist import shard_range
    for n in (0, 1, 5, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference: one JSON line with the contract's keys (tiny run)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "grids/s" and line["value"] > 0
    from oracle import ref_shim
    # the real reference wherever its tree is present (build container: /root/reference; GPU box: oracle/_ref)
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
