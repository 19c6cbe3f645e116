#!/usr/bin/env python
"""bench.py — SCENE-Net hot-path benchmark (BASELINE.json metric: voxel grids/s, 64^3 GENEO fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch: SceneNet forward (kernel synthesis + cast +
stencil/observer) and backward (tap-gradient reduction + parameter Jacobian) on a batch of 32
synthetic TS40K-shaped 64^3 occupancy grids per GPU (BASELINE config 2), driven by a
fixed upstream gradient dL/dpred ~ N(0,1).  Rank r owns its own batches (weak scaling); with
N > 1 the 13-float parameter-gradient payload is all-reduced (mean) over NCCL every step.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU = 32
GRID = (64, 64, 64)
KERNEL = (9, 5, 5)
P_OCC = 0.016
METRIC = "voxel grids/s (64^3 GENEO fwd+bwd)"
UNIT = "grids/s"
L2_BYTES = 126 * 1024 * 1024


# ----------------------------------------------------------------------------------------------
def kat_model(device, kernel=None):
    """SceneNet({'cy':1,'cone':1,'neg':1}, (9,5,5)) with the SURVEY §8c parameter vector."""
    import scenenet_b200 as sb
    params = {"cy_0.radius": 2.5, "cy_0.sigma": 1.8, "cone_0.apex": 4.0, "cone_0.cone_inc": 0.3, "cone_0.cone_radius": 2.0,
              "cone_0.radius": 3.0, "cone_0.sigma": 1.4, "neg_0.neg_factor": 0.2, "neg_0.radius": 8.0, "neg_0.sigma": 0.8}
    lambdas = {"lambda_cone_0": 0.3, "lambda_cy_0": 0.45, "lambda_neg_0": 0.25}
    torch.manual_seed(0)
    m = sb.SceneNet({'cy': 1, 'cone': 1, 'neg': 1}, tuple(kernel or KERNEL)).to(device)
    with torch.no_grad():
        for name, layer in m.geneos.items():
            for pn, p in layer.geneo_params.items():
                p.fill_(params[f"{name}.{pn}"])
        for ln, p in m.lambdas_dict.items():
            p.fill_(lambdas[ln])
            p.requires_grad_(ln != "lambda_cy_0")
    m.last_lambda = "lambda_cy_0"
    return m


def make_pool(device, rank, n_sets, dtype):
    """n_sets distinct (x, dpred) batches so that consecutive steps never find their inputs in L2."""
    pool = []
    for s in range(n_sets):
        g = torch.Generator(device=device).manual_seed(1234 + 1000 * rank + s)
        x = (torch.rand((B_PER_GPU, 1, *GRID), generator=g, device=device) < P_OCC).to(dtype)
        dp = torch.randn((B_PER_GPU, 1, *GRID), generator=g, device=device, dtype=torch.float32).to(dtype)
        pool.append((x, dp))
    return pool


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": (max(mx) if mx else None), "reasons": reasons,
                "samples": len(sm), "power_w_max": (max(pw) if pw else None)}


def voxel_bench(device, hbm_gbs):
    """BASELINE metric, second half: voxelization Mpts/s (config 3 shapes: UTM-offset float64 rows x,y,z,label;
    70 % ground sheet, 25 % vegetation blobs, 5 % tower column; scan-coherent order), counts + keep votes +
    occupancy/target grids, through voxel_ops.voxelize_clouds (bounding box, edges, binning, finalize)."""
    from scenenet_b200 import voxel_ops

    def cloud(n, seed):
        g = torch.Generator(device=device).manual_seed(seed)
        f64 = dict(generator=g, device=device, dtype=torch.float64)
        n_g, n_v = int(0.70 * n), int(0.25 * n)
        n_t = n - n_g - n_v
        ground = torch.stack([torch.rand(n_g, **f64) * 30, torch.rand(n_g, **f64) * 30, torch.randn(n_g, **f64) * 0.3], 1)
        cent = torch.rand((20, 3), **f64) * torch.tensor([30, 30, 9.0], device=device, dtype=torch.float64)
        veg = cent[torch.randint(0, 20, (n_v,), generator=g, device=device)] + torch.randn((n_v, 3), **f64) * 2.0
        tower = torch.stack([torch.randn(n_t, **f64) * 0.5 + 15, torch.randn(n_t, **f64) * 0.5 + 15, torch.rand(n_t, **f64) * 40], 1)
        pts = torch.cat([ground, veg, tower]) + torch.tensor([544850.0, 4634550.0, 160.0], device=device, dtype=torch.float64)
        lab = torch.cat([torch.randint(1, 13, (n_g + n_v,), generator=g, device=device).double(),
                         torch.full((n_t,), 15.0, device=device, dtype=torch.float64)])
        order = torch.argsort(((pts[:, 0] - 544850.0) / 2).floor() * 64 + ((pts[:, 1] - 4634550.0) / 2).floor())
        return torch.cat([pts[order], lab[order, None]], 1).contiguous()

    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps * 1e-3

    def graphed(fn):
        """the same call captured once in a CUDA graph and replayed (fixed shapes: one cudaGraphLaunch per cloud batch)"""
        s_ = torch.cuda.Stream(device=device)
        s_.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(s_):
            for _ in range(2):
                fn()
        torch.cuda.current_stream(device).wait_stream(s_)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()
        return g, keep

    out = {}
    for n, grid in ((1_000_000, 64), (10_000_000, 128)):
        rows = cloud(n, 1)
        call = lambda: voxel_ops.voxelize_clouds(rows[:, :3], None, (grid,) * 3, rows[:, 3], [15], want=("occ", "occ_keep"))
        dt = timeit(call)
        g, _keep = graphed(call)
        dtg = timeit(g.replay)
        alg = 56 * n + 24 * grid ** 3
        out[f"{n // 1_000_000}M_pts_{grid}^3"] = {"Mpts_per_s": n / dtg / 1e6, "us": dtg * 1e6, "algorithmic_GBps": alg / dtg / 1e9,
                                                   "hbm_frac": alg / dtg / 1e9 / hbm_gbs, "eager_us": dt * 1e6,
                                                   "eager_Mpts_per_s": n / dt / 1e6}
        del rows, g, _keep
    rows = torch.cat([cloud(60_000, s) for s in range(32)])
    off = torch.arange(0, 33, device=device, dtype=torch.int64) * 60_000
    call = lambda: voxel_ops.voxelize_clouds(rows[:, :3], off, (64, 64, 64), rows[:, 3], [15], want=("occ", "occ_keep"))
    dt = timeit(call)
    g, _keep = graphed(call)
    dtg = timeit(g.replay)
    alg = 56 * 32 * 60_000 + 24 * 32 * 64 ** 3
    out["batch_32x60k_pts_64^3"] = {"Mpts_per_s": 32 * 60_000 / dtg / 1e6, "clouds_per_s": 32 / dtg, "us": dtg * 1e6,
                                    "algorithmic_GBps": alg / dtg / 1e9, "hbm_frac": alg / dtg / 1e9 / hbm_gbs,
                                    "eager_us": dt * 1e6, "eager_clouds_per_s": 32 / dt}
    # the batched device loader (core/datasets/ts40k.py): 8 batches of 32 in-memory [60000, 4] float64 samples -> pinned staging
    # -> H2D on a copy stream -> one voxelize_clouds call per batch -> (x, y) float64 [32,1,64,64,64]; wall clock, host included
    try:
        import time
        from scenenet_b200 import TS40KDeviceLoader
        host_rows = [rows[i * 60_000:(i + 1) * 60_000].cpu().numpy() for i in range(32)] * 8
        loader = TS40KDeviceLoader(host_rows, batch_size=32, device=device)
        for _ in loader:  # warm-up epoch: pinned buffers allocated
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nb = 0
        for x_, y_ in loader:
            nb += x_.shape[0]
        torch.cuda.synchronize()
        dtl = time.perf_counter() - t0
        out["device_loader_32x60k_pts_64^3"] = {"clouds_per_s": nb / dtl, "Mpts_per_s": nb * 60_000 / dtl / 1e6, "h2d_bytes_per_batch": 32 * 60_000 * 32,
                                                "samples_page_locked_in_place": len(loader._registered),
                                                "note": "wall clock over 8 batches; the in-memory samples are page-locked in place (cudaHostRegister) and copied "
                                                        "to the device straight from the arrays, one cudaMemcpyAsync per cloud (PCIe bound); 0 locked samples = "
                                                        "the staging path (host memcpy into a pinned buffer, 12.8 k clouds/s)"}
    except Exception as e:  # noqa: BLE001
        out["device_loader_32x60k_pts_64^3"] = {"error": repr(e)}
    out["note"] = ("6 launches per call (bounding box, edges, grid init, binning, finalize), timed as CUDA-graph replays of the "
                   "captured call (eager figures beside them are host-bound); algorithmic bytes = 56 B/point + 24 B/voxel (SURVEY 8d)")
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU arms
CPU_ARM_SRC = r"""
import json, os, sys, time
sys.path.insert(0, {root!r})
import torch
assert not torch.cuda.is_available()
torch.set_num_threads(os.cpu_count() or 1)
from oracle import model_oracle as mo
batch, steps, warmup, kernel, grid, use_ref, budget_s = {batch}, {steps}, {warmup}, {kernel!r}, {grid!r}, {use_ref}, {budget_s}
x, _ = mo.synthetic_grids(batch, grid, seed=1234)
dpred = torch.randn(x.shape, generator=torch.Generator().manual_seed(1235), dtype=torch.float64)
if use_ref:
    # the UNMODIFIED reference: core/models/SCENE_Net.py SceneNet.forward (float64 conv3d of G kernels + observer) and
    # autograd's backward onto the 11 trainable scalars, on the host cores
    from oracle import ref_shim
    m = ref_shim.reference_scenenet(mo.KAT_GENEO_NUM, kernel, mo.KAT_PARAMS, mo.KAT_LAMBDAS, mo.KAT_LAST)
    def step():
        for p in m.parameters():
            p.grad = None
        m(x).backward(dpred)
else:
    model = mo.kat_model(kernel)
    def step():
        mo.fwd_bwd(model, x, None, dpred)
tw = time.perf_counter()
for i in range(warmup):
    step()
    if i >= 0 and time.perf_counter() - tw > 0.25 * budget_s:
        break
t0 = time.perf_counter()
done = 0
while done < steps:
    step()
    done += 1
    if done >= 2 and (time.perf_counter() - t0) * (done + 1) / done > budget_s:
        break  # bounded: the arm must end within a few minutes on any host
dt = time.perf_counter() - t0
print("CPU_ARM " + json.dumps(dict(grids_per_s=batch * done / dt, s_per_step=dt / done, threads=torch.get_num_threads(),
                                   kind="reference" if use_ref else "port", steps_timed=done)))
"""


def time_cpu(batch, steps, warmup, timeout=1500, budget_s=200.0):
    """fwd+bwd of config 2 on the host cores, in a child process that cannot see the GPU (the reference pins its
    kernels to CUDA whenever a GPU is visible, core/models/geneos/*.py): the REAL reference (`/root/reference` in the
    build container, its unmodified copy `oracle/_ref` on the GPU box — oracle/fetch_ref.py) when the tree is present,
    else the oracle port.  Returns (grids/s, s/step, threads, kind)."""
    from oracle import ref_shim
    src = CPU_ARM_SRC.format(root=ROOT, batch=batch, steps=steps, warmup=warmup, kernel=tuple(KERNEL), grid=tuple(GRID),
                             use_ref=bool(ref_shim.available()), budget_s=float(budget_s))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", PYTHONWARNINGS="ignore")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", src], env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    for ln in r.stdout.splitlines():
        if ln.startswith("CPU_ARM "):
            d = json.loads(ln[8:])
            return d["grids_per_s"], d["s_per_step"], d["threads"], d["kind"], d["steps_timed"]
    raise RuntimeError(f"CPU arm failed: {r.stdout[-1000:]} {r.stderr[-3000:]}")


def cpu_sample_note(batch, kind):
    what = ("the UNMODIFIED reference (core/models/SCENE_Net.py SceneNet fwd + autograd bwd, float64 conv3d, its copy oracle/_ref)"
            if kind == "reference" else "oracle port of the reference's float64 PyTorch CPU path (reference tree absent on this box)")
    whole = " (the whole config-2 batch)" if batch == B_PER_GPU else ""
    return f"{batch} of {B_PER_GPU} grids per step{whole}, {what}, all host threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch or B_PER_GPU  # the same config as the CUDA arm: 32 grids per step
    v, per_step, cores, kind, timed = time_cpu(batch, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
        "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "SCENE-Net training step fwd+bwd (13 params, 11 trainable), batch 32 of synthetic TS40K-shaped 64^3 "
                               "occupancy grids (Bernoulli 0.016), kernel (9,5,5), G=3, fixed upstream dL/dpred ~ N(0,1) (BASELINE "
                               "config 2); the reference's own CPU path on the host cores",
                   "global_batch": batch, "same_config_as_cuda_arm": batch == B_PER_GPU},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_note(batch, kind)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- config 5
def run_config5(args):
    rec = measure_config5(args.steps, args.warmup)
    if rec is not None:
        print(json.dumps(rec))
    sys.stdout.flush()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os._exit(0)


def measure_config5(steps, warmup):
    """BASELINE config 5: SemanticKITTI-shaped synthetic scans (~120 k points, float64 rows x,y,z,label; ring pattern
    in +-50 m, z in [-3, 3], pole label 80 kept) -> voxelize to (64, 64, 256) -> SceneNet inference -> threshold 0.65.
    One rank per GPU, SCANS scans per step per GPU; e2e includes the H2D copy of the points and the D2H of the
    per-scan positive-voxel counts."""
    import scenenet_b200 as sb
    from scenenet_b200 import dist as sdist, voxel_ops
    import torch.distributed as dist
    rank, world, device = sdist.init_from_env()
    SCANS, NPTS, GRID_XYZ = 8, 120_000, (64, 64, 256)
    model = kat_model(device)
    g = torch.Generator().manual_seed(77 + rank)

    def scans():
        r = 5.0 + 45.0 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64) ** 2      # range-weighted rings
        th = 2 * 3.141592653589793 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64)
        z = -3.0 + 6.0 * torch.rand(SCANS * NPTS, generator=g, dtype=torch.float64) ** 3
        lab = torch.where(torch.rand(SCANS * NPTS, generator=g) < 0.01, 80.0, 40.0).to(torch.float64)
        pts = torch.stack([(r * torch.cos(th)).float().double(), (r * torch.sin(th)).float().double(), z.float().double(), lab], 1)
        return pts.contiguous().pin_memory()

    host = [scans() for _ in range(2)]
    off = (torch.arange(0, SCANS + 1, dtype=torch.int64) * NPTS).to(device)
    dev_rows = [torch.empty_like(h, device=device) for h in host]
    counts_host = [torch.zeros(SCANS, dtype=torch.int64).pin_memory() for _ in range(2)]

    def infer(rows):
        out = voxel_ops.voxelize_clouds(rows[:, :3], off, GRID_XYZ, rows[:, 3], [80.0], want=("occ",), occ_dtype=torch.float32)
        nz, nx, ny = GRID_XYZ[2], GRID_XYZ[0], GRID_XYZ[1]
        with torch.no_grad():
            pred = model(out["occ"].view(SCANS, 1, nz, nx, ny))
        lab = sb.voxelization.prob_to_label(pred, 0.65)
        return lab.view(SCANS, -1).sum(1).to(torch.int64)

    # resident-input value: CUDA-graph replay of voxelize + forward + threshold
    for j in range(2):
        dev_rows[j].copy_(host[j])
    graphs, outs = [], []
    for j in range(2):
        s_ = torch.cuda.Stream(device=device)
        s_.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(s_):
            for _ in range(2):
                infer(dev_rows[j])
        torch.cuda.current_stream(device).wait_stream(s_)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            o = infer(dev_rows[j])
        graphs.append(gr)
        outs.append(o)

    def barrier():
        if world > 1:
            dist.barrier()

    for i in range(warmup):
        graphs[i % 2].replay()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % 2].replay()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    value = SCANS * world * steps / (float(ms) * 1e-3)
    # e2e: points from pinned host memory every step, per-scan counts back to the host
    t0 = time.perf_counter()
    n_e2e = max(3, min(steps, 20))
    for i in range(n_e2e):
        j = i % 2
        dev_rows[j].copy_(host[j], non_blocking=True)
        graphs[j].replay()
        counts_host[j].copy_(outs[j], non_blocking=True)
        torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = SCANS * world * n_e2e / float(dt)
    if rank == 0:
        return ({
            "metric": "scans/s (voxelize + GENEO inference, KITTI-shaped)", "value": value, "unit": "scans/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": float(ms) / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{SCANS} SemanticKITTI-shaped scans x {NPTS} points per GPU -> (64,64,256) occupancy grids -> "
                                   "SceneNet (9,5,5) forward -> threshold 0.65 (BASELINE config 5)", "parallelism": f"dp{world}",
                       "launch": "CUDA-graph replay"},
            "Mpts_per_s": value * NPTS / 1e6,
            "e2e": {"value": e2e, "unit": "scans/s", "h2d_bytes_per_step": host[0].numel() * 8, "d2h_bytes_per_step": SCANS * 8},
            "positive_voxels_first_scan": int(outs[0][0]), "clocks": clocks})
    return None


def measure_config4(device, rank, world, sync_group, kernel=9, steps=10, warmup=3, batch=8):
    """BASELINE config 4 in short: 128^3 grids, cubic kernel^3 GENEO kernels, `batch` grids per GPU, fwd + bwd with a fixed
    upstream gradient, the parameter gradients summed over the ranks inside backward (weak scaling).  CUDA-graph replay."""
    import torch.distributed as dist
    from scenenet_b200.graphs import GraphedStep
    model = kat_model(device, (kernel,) * 3)
    model.grad_scale = 1.0 / world
    model.grad_sync_group = sync_group
    pool = []
    for s_ in range(2):
        g = torch.Generator(device=device).manual_seed(1234 + 1000 * rank + s_)
        x = (torch.rand((batch, 1, 128, 128, 128), generator=g, device=device) < P_OCC).to(torch.float64)
        dp = torch.randn((batch, 1, 128, 128, 128), generator=g, device=device, dtype=torch.float32).to(torch.float64)
        pool.append((x, dp))
    graphs = [GraphedStep(model, x, dpred=dp, specialize=True) for x, dp in pool]
    for i in range(warmup):
        graphs[i % 2].replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        graphs[i % 2].replay()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    del graphs, pool
    return {"metric": f"voxel grids/s (128^3 GENEO fwd+bwd, {kernel}^3 kernels)", "value": batch * world * steps / (float(ms) * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "ms_per_step": float(ms) / steps, "scaling": "weak",
            "config": {"workload": f"batch {batch} per GPU of synthetic 128^3 occupancy grids (Bernoulli {P_OCC}), kernel ({kernel},)*3, G=3 "
                                   "(BASELINE config 4)", "parallelism": f"dp{world}"}}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--io-dtype", default="f64", choices=["f64", "f32"], help="dtype of x / dpred / pred at the module boundary")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=0, help="grids per CPU step of the reference arm (default: the whole batch of 32; "
                                                             "smaller values are for quick self-tests and are reported as a sample)")
    ap.add_argument("--eager", action="store_true", help="time the eager module path instead of CUDA-graph replays")
    ap.add_argument("--workload", default="config2", choices=["config2", "config4", "config5"],
                    help="config2 = the headline (default); config4 = 128^3 grids, cubic kernel --kernel, batch 8 per GPU; "
                         "config5 = KITTI-shaped scans: voxelize + GENEO inference")
    ap.add_argument("--kernel", type=int, default=9, help="config4: cubic kernel extent (9, 11, 13, 15)")
    ap.add_argument("--no-sub-records", action="store_true", help="N > 1: do not add the short config 4 / config 5 records")
    ap.add_argument("--brief", action="store_true", help="skip the rank-0-only kernel roofline / voxelization / CPU sections")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "config5":
        return run_config5(args)
    if args.workload == "config4":
        global B_PER_GPU, GRID, KERNEL, METRIC
        B_PER_GPU, GRID, KERNEL = 8, (128, 128, 128), (args.kernel,) * 3
        METRIC = f"voxel grids/s (128^3 GENEO fwd+bwd, {args.kernel}^3 kernels)"
        args.no_cpu_baseline = True

    import scenenet_b200 as sb
    from scenenet_b200 import dist as sdist, ops
    import torch.distributed as dist

    rank, world, device = sdist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scenenet_b200 hot path has no CPU fallback")
    io_dtype = torch.float64 if args.io_dtype == "f64" else torch.float32
    model = kat_model(device)
    trainable = [p for p in model.parameters() if p.requires_grad]
    model.grad_scale = 1.0 / world  # DDP mean folded into the Jacobian kernel; the collective only sums
    grad_allreduce = None
    if world > 1:
        # one all-reduce of the flat 13-float payload inside backward: our single-kernel exchange over NVLink peer
        # memory when the peer mappings are available, else NCCL
        try:
            if os.environ.get("SN_BENCH_NCCL"):
                raise RuntimeError("forced by SN_BENCH_NCCL")
            model.grad_sync_group = sdist.PeerAllReduce(device)
            grad_allreduce = ("exchange over NVLink peer memory fused into the parameter-Jacobian kernel "
                              "(sn_scenenet_param_grads_allreduce: no separate collective launch)")
        except Exception as e:  # noqa: BLE001
            model.grad_sync_group = True
            grad_allreduce = f"ncclAllReduce (peer memory unavailable: {type(e).__name__}: {str(e)[:80]})"
        # every rank must take the same path
        flag = torch.tensor([1 if callable(model.grad_sync_group) else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 0 and callable(model.grad_sync_group):
            model.grad_sync_group = True
            grad_allreduce = "ncclAllReduce (peer memory unavailable on another rank)"

    bytes_per_set = B_PER_GPU * GRID[0] * GRID[1] * GRID[2] * (8 if io_dtype == torch.float64 else 4) * 3
    n_sets = max(3, -(-4 * L2_BYTES // bytes_per_set))
    pool = make_pool(device, rank, n_sets, io_dtype)

    from scenenet_b200.graphs import GraphedStep
    post = None  # the gradient all-reduce lives inside the model's backward (model.grad_sync_group)

    def eager_step(i):
        x, dp = pool[i % n_sets]
        for p in trainable:
            p.grad = None
        pred = model(x)
        pred.backward(dp)
        if post:
            post()
        return pred

    graphs = None
    if not args.eager:
        # one captured step per input set: static addresses, no staging copies inside the timed region
        graphs = [GraphedStep(model, x, dpred=dp, post_backward=post, specialize=True) for x, dp in pool]

    def step(i):
        if graphs is None:
            return eager_step(i)
        return graphs[i % n_sets].replay()

    def barrier():
        if world > 1:
            dist.barrier()

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    barrier()

    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    n0 = sb._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    launches = sb._lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    value = B_PER_GPU * world * args.steps / (total_ms * 1e-3)
    # synthesis, prepare (cast + non-zero count + occupancy bits), forward, g0, tap gradient, row sum, parameter Jacobian^T
    # (which also exchanges the gradients over NVLink when N > 1); specialised capture: no gated-out
    # launches.  Replayed graph nodes: the library's host-side counter saw them once, at capture (GraphedStep.kernels_per_replay)
    if graphs is not None:
        launches = sum(graphs[(args.warmup + i) % n_sets].kernels_per_replay for i in range(args.steps))

    # eager module path (what a drop-in user gets without graph capture), a few steps, for the record
    for i in range(3):
        eager_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        eager_step(i)
    torch.cuda.synchronize()
    eager_value = B_PER_GPU * world * 10 / (time.perf_counter() - t0)

    # the same step fed with the grids as ONE BIT PER VOXEL, device-resident (what TS40KDeviceLoader(dtype=torch.int32) hands the
    # model: the float64 -> float32 + bits preparation pass, 20.7 us of re-formatting, becomes an 8 us expansion of the bits)
    packed_value = None
    if graphs is not None and args.workload == "config2" and not args.brief:
        pk = [(ops.pack_occupancy(x), dp.float()) for x, dp in pool]
        pgraphs = [GraphedStep(model, xb, dpred=dpf, post_backward=post, specialize=True) for xb, dpf in pk]
        for i in range(args.warmup):
            pgraphs[i % n_sets].replay()
        torch.cuda.synchronize()
        barrier()
        pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pa.record()
        for i in range(args.steps):
            pgraphs[(args.warmup + i) % n_sets].replay()
        pb.record()
        torch.cuda.synchronize()
        pms = torch.tensor([pa.elapsed_time(pb)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(pms, op=dist.ReduceOp.MAX)
        packed_value = {"value": B_PER_GPU * world * args.steps / (float(pms) * 1e-3), "unit": UNIT, "ms_per_step": float(pms) / args.steps,
                        "note": "device-resident input as one bit per voxel (ops.pack_occupancy), float32 pred / dpred: same step, "
                                "same kernels behind the preparation; predictions and gradients bit-identical to the float32-input step"}
        del pgraphs, pk

    # ------------------------------------------------ e2e: host buffers, copies inside the timed region
    n_e2e = max(3, min(args.steps, 20))
    gh = torch.zeros(len(trainable), dtype=torch.float32).pin_memory()

    def e2e_measure(xh, dps):
        """xh: two pinned host batches; dps: two device upstream gradients (dtype of the module's output).
        Every step: H2D of that step's x (copy stream, double-buffered, overlapping the previous step's kernels) ->
        fwd + bwd [+ all-reduce] + D2H of the gradients -> the host waits for the gradients."""
        xd2 = [torch.empty(xh[0].shape, dtype=xh[0].dtype, device=device) for _ in range(2)]
        gh2 = [torch.zeros(len(trainable), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=device)
        main_stream = torch.cuda.current_stream(device)
        ev_copied = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        for e in ev_done:
            e.record(main_stream)
        if args.eager:
            steps_ = None
        else:
            steps_ = [GraphedStep(model, xd2[j], dpred=dps[j], grads_host=gh2[j], post_backward=post) for j in range(2)]
        state = {"next": 0}

        def prefetch(i):
            j = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_done[j])          # the step that last used this buffer has finished
                xd2[j].copy_(xh[j], non_blocking=True)      # H2D of step i's input grids
                ev_copied[j].record(copy_stream)

        def one(i):
            j = i % 2
            while state["next"] <= i + 1:                    # this step's input, and the next one's while we compute
                prefetch(state["next"])
                state["next"] += 1
            main_stream.wait_event(ev_copied[j])
            if steps_ is not None:
                steps_[j].replay()
            else:
                for p in trainable:
                    p.grad = None
                pred = model(xd2[j])
                pred.backward(dps[j])
                if post:
                    post()
                gh2[j].copy_(torch.stack([p.grad for p in trainable]), non_blocking=True)
            ev_done[j].record(main_stream)
            ev_done[j].synchronize()                         # the host reads this step's gradients
            gh.copy_(gh2[j])

        # warm-up: at least 3 steps AND 50 ms — after the PCIe-bound float64 phase the GPU sits nearly idle and the first
        # milliseconds of the next phase ran at half speed on some boxes (uint8 input: 66-71 k against 135-168 k grids/s)
        # (a step COUNT, the same on every rank: the steps contain the gradient exchange)
        est = max(xh[0].numel() * xh[0].element_size() / 50e9, 1.5e-4)
        i0w = max(3, min(400, int(0.05 / est)))
        for i in range(i0w):
            one(i)
        # wall clock over n_e2e steps, three times; the MEDIAN is reported (one host hiccup inside a 5 ms window halves a
        # single reading: a run measured 66 k and 166 k grids/s for the same uint8 path on two boxes)
        reads, i0 = [], i0w
        for _rep in range(3):
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(i0, i0 + n_e2e):
                one(i)
            torch.cuda.synchronize()
            dt_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
            reads.append(float(dt_))
            i0 += n_e2e
        dt = sorted(reads)[1]
        return B_PER_GPU * world * n_e2e / float(dt), xh[0].numel() * xh[0].element_size()

    xh = [torch.empty(pool[0][0].shape, dtype=io_dtype).pin_memory() for _ in range(2)]
    for j, h in enumerate(xh):
        h.copy_(pool[j][0])
    # host->device bandwidth of this box for one batch from pinned memory: the ceiling of the float64 e2e figure
    xd_probe = torch.empty_like(pool[0][0])
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xd_probe.copy_(xh[0], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    ev0.record()
    for _ in range(5):
        xd_probe.copy_(xh[0], non_blocking=True)
    ev1.record()
    ev1.synchronize()
    h2d_gbps = 5 * xh[0].numel() * xh[0].element_size() / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del xd_probe
    e2e_value, h2d = e2e_measure(xh, [pool[0][1], pool[1][1]])
    d2h = gh.numel() * gh.element_size()
    # the same step fed with byte occupancy grids (what ToFullDense produces; module extension): 8x fewer PCIe bytes
    xh8 = [h.to(torch.uint8).pin_memory() for h in xh]
    e2e_u8_value, h2d_u8 = e2e_measure(xh8, [pool[0][1].float(), pool[1][1].float()])
    # ... and with one BIT per voxel (ops.pack_occupancy, SN_BITS): 64x fewer bytes than the reference's float64 grids
    xhb = [ops.pack_occupancy(h).pin_memory() for h in xh]
    e2e_bits_value, h2d_bits = e2e_measure(xhb, [pool[0][1].float(), pool[1][1].float()])

    # ------------------------------------------------ config 2 (ii): full training_step semantics — the module step with the
    # drop-in GENEO_Tversky_Loss (fused criterion kernels) instead of a fixed upstream gradient
    train_value = None
    if world == 1 and args.workload == "config2":
        HIST = ([52648, 52727, 52553, 52392, 52366, 52380, 52501, 51922, 52499, 52300], [0.1 * k for k in range(10)])
        crit = sb.GENEO_Tversky_Loss(hist=HIST, weight_alpha=1, weight_epsilon=0.1, mse_weight=1, convex_weight=5, tversky_alpha=2,
                                     tversky_beta=1, focal_gamma=4, tversky_smooth=1e-6)
        ys = []
        for s_ in range(n_sets):
            g = torch.Generator(device=device).manual_seed(4321 + s_)
            ys.append((torch.rand((B_PER_GPU, 1, *GRID), generator=g, device=device) < 3e-4).to(io_dtype))
        tgraphs = [GraphedStep(model, pool[s_][0], loss_fn=(lambda pred, y=ys[s_]: crit(pred, y, model.get_cvx_coefficients(),
                                                                                         model.get_geneo_params())), specialize=True)
                   for s_ in range(n_sets)]
        for i in range(args.warmup):
            tgraphs[i % n_sets].replay()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(args.steps):
            tgraphs[(args.warmup + i) % n_sets].replay()
        b_.record()
        b_.synchronize()
        train_ms = a_.elapsed_time(b_) / args.steps
        # the same step with the observer and the criterion as ONE autograd node (criterion.training_loss: G0 comes
        # straight from the criterion's backward; an extension next to the drop-in path)
        sgraphs = [GraphedStep(model, pool[s_][0], loss_from_input=(lambda x_, y=ys[s_]: crit.training_loss(model, x_, y)),
                               specialize=True) for s_ in range(n_sets)]
        for i in range(args.warmup):
            sgraphs[i % n_sets].replay()
        torch.cuda.synchronize()
        a_.record()
        for i in range(args.steps):
            sgraphs[(args.warmup + i) % n_sets].replay()
        b_.record()
        b_.synchronize()
        single_ms = a_.elapsed_time(b_) / args.steps
        single_loss = float(sgraphs[0].loss.detach())
        del sgraphs
        train_value = {"value": B_PER_GPU / (train_ms * 1e-3), "unit": UNIT, "ms_per_step": train_ms,
                       "loss": float(tgraphs[0].loss.detach()),
                       "single_node": {"value": B_PER_GPU / (single_ms * 1e-3), "unit": UNIT, "ms_per_step": single_ms, "loss": single_loss,
                                       "note": "criterion.training_loss(model, x, y): no dL/dpred tensor, no separate G0 pass"},
                       "note": "config 2(ii): forward + GENEO_Tversky_Loss (drop-in class, fused reduction / closed-form "
                               "dL/dpred kernels, penalties on the live parameters) + backward, CUDA-graph replay"}
        del tgraphs

    grad_sync_ok = None
    if world > 1:
        # after the all-reduce every rank must hold the same (mean) gradients
        step(0)
        torch.cuda.synchronize()
        flat = torch.stack([p.grad.reshape(()) for p in trainable])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        grad_sync_ok = bool(all(torch.equal(g, gathered[0]) for g in gathered)) and bool(flat.abs().sum() > 0)
        if callable(model.grad_sync_group):
            # the peer-memory exchange against NCCL on the same per-rank payload
            grad_sync_ok = grad_sync_ok and model.grad_sync_group.ok()
            probe = torch.arange(1, 14, device=device, dtype=torch.float32) * (rank + 1) * 0.37
            ref = probe.clone()
            dist.all_reduce(ref, op=dist.ReduceOp.SUM)
            model.grad_sync_group(probe)
            torch.cuda.synchronize()
            grad_sync_ok = grad_sync_ok and bool(torch.allclose(probe, ref, rtol=1e-6, atol=0))
    barrier()
    # BASELINE configs 4 and 5 in short next to the headline whenever several GPUs are used (so that the driver's scaling
    # runs record them at every N): a few steps each, all ranks take part
    sub_records = None
    if world > 1 and args.workload == "config2" and not args.no_sub_records:
        sub_records = {}
        for name, fn in (("config4_9^3", lambda: measure_config4(device, rank, world, model.grad_sync_group, 9)),
                         ("config4_15^3", lambda: measure_config4(device, rank, world, model.grad_sync_group, 15, steps=4, warmup=2)),
                         ("config5", lambda: measure_config5(10, 3))):
            try:
                sub_records[name] = fn()
            except Exception as e:  # noqa: BLE001
                sub_records[name] = {"error": repr(e)[:200]}
            barrier()
    if world > 1 and rank != 0:
        sys.stdout.flush()
        os._exit(0)  # ranks > 0 are done: the remaining sections are rank-0-only and use no collective
    # ------------------------------------------------ roofline of the two stencil kernels, timed alone with CUDA events
    roof = None
    cpu_base = None
    vox = None
    if rank == 0 and not args.brief:
        T = KERNEL[0] * KERNEL[1] * KERNEL[2]
        V = B_PER_GPU * GRID[0] * GRID[1] * GRID[2]
        peak_tf = ops.fp32_peak_probe(2000, device)
        hbm_gbs, hbm_src = measured_peaks()
        from scenenet_b200._lib import SN_TAPGRAD_DENSE, SN_TAPGRAD_SPARSE
        prep = [ops.prepare(p[0]) for p in pool]
        x32s = [a for a, _ in prep]
        K, lam, Kstar, snap = ops.synth_fwd(*_spec_params(model))
        preds = [ops.scenenet_fwd(x32, Kstar, io_dtype) for x32 in x32s]
        reps = 20

        def time_kernel(fn):
            """average duration of one call of fn (one kernel of this library), CUDA events on the launching stream around
            replays of a CUDA graph that holds one call per input set (no host gaps between the launches: the eager loop
            used in round 1 was host-bound below ~30 us); inputs rotate over the > L2 pool"""
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            try:
                s_ = torch.cuda.Stream(device=device)
                s_.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(s_):
                    fn(0)
                torch.cuda.current_stream(device).wait_stream(s_)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    keep = [fn(i) for i in range(n_sets)]
                gr.replay()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                nrep = max(2, reps // n_sets)
                a.record()
                for _ in range(nrep):
                    gr.replay()
                b.record()
                b.synchronize()
                del keep
                return a.elapsed_time(b) / (nrep * n_sets) * 1e-3
            except Exception:  # noqa: BLE001  (capture unavailable: eager loop)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(reps):
                    fn(i)
                b.record()
                b.synchronize()
                return a.elapsed_time(b) / reps * 1e-3

        g0s = [ops.g0(preds[i], pool[i][1]) for i in range(n_sets)]
        from scenenet_b200._lib import SN_PATH_DENSE, SN_PATH_SPARSE
        t_fwd = time_kernel(lambda i: ops.scenenet_fwd(x32s[i % n_sets], Kstar, io_dtype, mode=SN_PATH_DENSE))
        try:  # the occupancy-driven forward, mask-driven from sn_grid_prepare's occupancy bits (what a step runs)
            t_fwd_sp = time_kernel(lambda i: ops.scenenet_fwd(x32s[i % n_sets], Kstar, io_dtype, nnz=prep[i % n_sets][1], mode=SN_PATH_SPARSE))
        except Exception:  # noqa: BLE001  (slices of more than 128 taps: no occupancy-driven forward, the dense stencil runs)
            t_fwd_sp = None
        t_fwd_auto = time_kernel(lambda i: ops.scenenet_fwd(x32s[i % n_sets], Kstar, io_dtype, nnz=prep[i % n_sets][1]))
        t_g0 = time_kernel(lambda i: ops.g0(preds[i % n_sets], pool[i % n_sets][1]))
        # tap gradient (+ row reduction): the dense stencil, the occupancy-driven kernel, and what a step runs
        # (both enqueued, the device picks one from the non-zero count)
        t_tap = time_kernel(lambda i: ops.tapgrad(x32s[i % n_sets], g0s[i % n_sets], KERNEL, mode=SN_TAPGRAD_DENSE))
        t_tap_sp = time_kernel(lambda i: ops.tapgrad(x32s[i % n_sets], g0s[i % n_sets], KERNEL, mode=SN_TAPGRAD_SPARSE))
        t_tap_auto = time_kernel(lambda i: ops.tapgrad(x32s[i % n_sets], g0s[i % n_sets], KERNEL, nnz=prep[i % n_sets][1]))
        t_cast = time_kernel(lambda i: ops.prepare(pool[i % n_sets][0]))
        fl = 2.0 * T * V
        esz = 8 if io_dtype == torch.float64 else 4
        # the dominant kernel of a step is the forward.  Which forward runs is decided on the device from the occupancy:
        # the dense stencil is bound by the FP32 pipe (2 T flop per voxel); the occupancy-driven kernel executes only the
        # multiply-adds of the occupied voxels, so it is held to the HBM roofline of its algorithmic bytes (x read once in
        # float32 + pred written once) and the dense stencil's FP32 figures are reported beside it, clearly labelled.
        nnz0 = int(prep[0][1][0])
        fwd_sparse_runs = t_fwd_sp is not None and ops.lib.sn_select_path(0, nnz0, B_PER_GPU, *GRID, *KERNEL) == SN_PATH_SPARSE
        bytes_fwd = V * (4 + esz)
        g0_bytes = V * (2 * esz + 4)
        dense_fp32 = {"us": t_fwd * 1e6, "tflops": fl / t_fwd / 1e12, "frac": fl / t_fwd / 1e12 / peak_tf}
        if fwd_sparse_runs:
            dom, t_dom = "fwd_occ_kernel", t_fwd_sp
            roof = {"bound": "hbm", "kernel": dom, "achieved": bytes_fwd / t_dom / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                    "frac": bytes_fwd / t_dom / 1e9 / hbm_gbs, "traffic": _ncu_traffic(dom), "peak_source": hbm_src,
                    "algorithmic": {"bytes_per_voxel": 4 + esz, "voxels_per_launch": V, "bytes_per_launch": bytes_fwd},
                    "note": "occupancy-driven forward (selected on the device at %.2f %% occupancy): cost follows the occupied "
                            "voxels (shared-memory scatter + instruction issue), not the 2 T flop per voxel of the dense "
                            "formulation; dense_equivalent_tflops counts flops it does NOT execute and is no roofline claim"
                            % (100.0 * nnz0 / V),
                    "dense_equivalent_tflops": fl / t_dom / 1e12}
        else:
            dom, t_dom = "stencil_fwd_kernel", t_fwd
            roof = {"bound": "fp32", "kernel": dom, "achieved": fl / t_dom / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": fl / t_dom / 1e12 / peak_tf, "traffic": _ncu_traffic(dom),
                    "peak_source": "FP32 FFMA probe measured in this run (sn_fp32_peak_probe); MEASURED_PEAKS.json has no FP32-pipe figure",
                    "algorithmic": {"flops_per_voxel_per_kernel": 2 * T, "voxels_per_launch": V, "bytes_per_launch": bytes_fwd},
                    "hbm": {"achieved": bytes_fwd / t_dom / 1e9, "peak": hbm_gbs, "unit": "GB/s", "frac": bytes_fwd / t_dom / 1e9 / hbm_gbs,
                            "peak_source": hbm_src}}
        roof.update({
            "fp32_peak_tflops": peak_tf,
            "fwd_dense": dict(dense_fp32, bound="fp32", note="direct stencil, TMA-staged halo, 8 x 4 register block; FP32 FFMA roofline"),
            "fwd_occupancy_driven": {"us": None if t_fwd_sp is None else t_fwd_sp * 1e6, "us_auto_selected": t_fwd_auto * 1e6,
                                     "GBps": None if t_fwd_sp is None else bytes_fwd / t_fwd_sp / 1e9,
                                     "hbm_frac": None if t_fwd_sp is None else bytes_fwd / t_fwd_sp / 1e9 / hbm_gbs,
                                     "note": "non-zero voxels listed from sn_grid_prepare's occupancy bits (one block-wide scan per tile), "
                                             "scattered into shared-memory planes, table-driven float64 tanh; us_auto_selected = the "
                                             "per-tile choice (ABI v4: tiles above ~5.5 % occupancy go to the dense stencil's tile-list "
                                             "pass, which finds nothing to do on these uniform grids)"},
            "bwd_tapgrad_dense": {"us": t_tap * 1e6, "tflops": fl / t_tap / 1e12, "frac": fl / t_tap / 1e12 / peak_tf},
            "bwd_tapgrad_occupancy_driven": {"us": t_tap_sp * 1e6, "us_auto_selected": t_tap_auto * 1e6, "bound": "hbm",
                                             "GBps": V * 8 / t_tap_sp / 1e9, "hbm_frac": V * 8 / t_tap_sp / 1e9 / hbm_gbs,
                                             "with_state_buffer": {"us": t_tap_auto * 1e6, "GBps": V * 4.125 / t_tap_auto / 1e9,
                                                                   "hbm_frac": V * 4.125 / t_tap_auto / 1e9 / hbm_gbs,
                                                                   "note": "binary grids with the grid state (what the step runs): voxels from "
                                                                           "the occupancy bits, G0 4 B + 1 bit per voxel, + the row-sum kernel"},
                                             "note": "without the state buffer: x + G0 read once (8 B/voxel), + the row-sum kernel; skips the 98.4 % zero voxels, so the FP32 "
                                                     "roofline of the dense formulation does not apply"},
            "g0_pass": {"us": t_g0 * 1e6, "GBps": g0_bytes / t_g0 / 1e9, "hbm_frac": g0_bytes / t_g0 / 1e9 / hbm_gbs},
            "prepare_pass": {"us": t_cast * 1e6, "GBps": V * (esz + (4 if esz == 8 else 0)) / t_cast / 1e9,
                             "hbm_frac": V * (esz + (4 if esz == 8 else 0)) / t_cast / 1e9 / hbm_gbs},
            "step_roofline_grids_per_s": {"fp32_dense_formulation": B_PER_GPU / (2 * fl / (peak_tf * 1e12)),
                                          "hbm_20B_per_voxel": B_PER_GPU / (20.0 * V / (hbm_gbs * 1e9))},
        })
        vox = voxel_bench(device, hbm_gbs) if args.workload == "config2" else None
        if not args.no_cpu_baseline and world == 1:
            v, per, cores, kind, _ = time_cpu(B_PER_GPU, 4, 1, budget_s=45.0)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "s_per_step": per,
                        "sample": cpu_sample_note(B_PER_GPU, kind) + "; 1 warm-up + up to 4 timed steps (at most ~45 s)"}

    if rank == 0:
        if roof is not None:
            # t_roof = max(4 T V B / P_fp32, 20 V B / BW_hbm) for one fwd + bwd step of the minimal-work dense
            # formulation; the step reaches a larger share of it than the dense stencils could (0.55 / 0.52) because the
            # occupancy-driven kernels do not execute the multiply-adds of empty voxels
            t_step = total_ms / args.steps * 1e-3
            t_fp32, t_hbm = 2 * fl / (roof["fp32_peak_tflops"] * 1e12), 20.0 * V / (hbm_gbs * 1e9)
            roof["step_vs_survey_8d_roofline"] = {"t_roof_us": max(t_fp32, t_hbm) * 1e6, "t_fp32_us": t_fp32 * 1e6, "t_hbm_us": t_hbm * 1e6,
                                                  "t_measured_us": t_step * 1e6, "frac": max(t_fp32, t_hbm) / t_step}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("SCENE-Net training step fwd+bwd (13 params, 11 trainable), batch 32 per GPU of synthetic "
                                    "TS40K-shaped 64^3 occupancy grids (Bernoulli 0.016), kernel (9,5,5), G=3, fixed upstream "
                                    "dL/dpred ~ N(0,1) (BASELINE config 2)") if args.workload == "config2" else
                                   (f"SCENE-Net fwd+bwd, batch {B_PER_GPU} per GPU of synthetic {GRID[0]}^3 occupancy grids (Bernoulli "
                                    f"0.016), kernel {KERNEL}, G=3, fixed upstream gradient (BASELINE config 4)"),
                       "global_batch": B_PER_GPU * world, "io_dtype": args.io_dtype, "parallelism": f"dp{world}", "grad_allreduce": grad_allreduce,
                       "launch": "eager module calls" if args.eager else "CUDA-graph replay of the captured module step (scenenet_b200.graphs.GraphedStep)",
                       "eager_module_value": eager_value, "packed_input_value": packed_value,
                       "l2": f"inputs rotate over {n_sets} distinct batches ({n_sets * bytes_per_set / 2**20:.0f} MiB) > 126 MiB L2; no flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e, "readings": "median of 3 x steps",
                    "uint8_occupancy_input": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8},
                    "packed_occupancy_input": {"value": e2e_bits_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bits,
                                               "note": "one bit per voxel (ops.pack_occupancy / TS40KDeviceLoader(dtype=torch.int32)); "
                                                       "same predictions and gradients as the float32 path, bit for bit"},
                    "h2d_GBps_measured": h2d_gbps, "pcie_bound": B_PER_GPU * world / (h2d / (h2d_gbps * 1e9)),
                    "pcie_note": "pcie_bound = grids/s at which the H2D copy of x alone saturates the measured host->device rate "
                                 "(all ranks copying at once)",
                    "note": "x (module-boundary dtype) from pinned host memory every step, double-buffered on a copy stream; "
                            "dL/dpred resident on the device (config 2(i)); gradients copied back and read by the host every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base, "training_step": train_value, "voxelize": vox, "grad_sync_ok": grad_sync_ok,
            "other_configs": sub_records,
        }
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        # no destroy_process_group(): tearing NCCL down under live CUDA graphs that contain collectives hung the
        # first 2-GPU run after the result line had been printed; every rank leaves once rank 0 is done printing
        os._exit(0)


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full capture
    (profiles/r2_traffic.json), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            d = json.load(f)[kernel]
        return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])
    except Exception:  # noqa: BLE001
        return None


def _spec_params(model):
    spec, params = model._spec_and_params()
    return spec, [p.detach() for p in params]


if __name__ == "__main__":
    main()
