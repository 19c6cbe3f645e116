"""Two GPUs: the single-kernel all-reduce over NVLink peer memory (csrc/peer_allreduce.cu) against NCCL, eagerly
and replayed from a CUDA graph.  Skipped on a one-GPU box."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from scenenet_b200 import dist as sdist
    r, w, dev = sdist.init_from_env()
    try:
        ar = sdist.PeerAllReduce(dev)
    except Exception as e:  # noqa: BLE001
        q.put((rank, "unavailable", f"{type(e).__name__}: {e}"))
        return
    ok = True
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(50):
        n = 1 + (it * 7) % 96
        v = torch.randn(n, generator=g, device=dev)
        ref = v.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ar(v)
        ok = ok and bool(torch.allclose(v, ref, rtol=1e-6, atol=1e-6))
    # CUDA-graph replay: the kernel keeps its own sequence number
    buf = torch.zeros(13, device=dev)
    src = torch.zeros(13, device=dev)
    s_ = torch.cuda.Stream(device=dev)
    s_.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s_):
        buf.copy_(src); ar(buf)
    torch.cuda.current_stream(dev).wait_stream(s_)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        ar(buf)
    for it in range(20):
        src.copy_(torch.arange(13, device=dev, dtype=torch.float32) * (rank + 1) + it)
        graph.replay()
        torch.cuda.synchronize()
        expect = torch.arange(13, device=dev, dtype=torch.float32) * sum(range(1, world + 1)) + it * world
        ok = ok and bool(torch.equal(buf, expect))
    ok = ok and ar.ok()
    dist.barrier()
    q.put((rank, "ok" if ok else "mismatch", ""))


def test_peer_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    if any(r[1] == "unavailable" for r in res):
        pytest.skip(f"peer memory unavailable: {res}")
    assert all(r[1] == "ok" for r in res), res
