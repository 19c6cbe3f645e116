"""Tensor-level wrappers over the C ABI (include/scenenet_b200.h).

torch is plumbing here: it owns device memory and the stream; all arithmetic happens in the
hand-written sm_100a kernels of libscenenet_b200.so.  Every function requires CUDA tensors
and raises otherwise — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import SN_BITS, SN_F32, SN_F64, SN_I32, SN_I64, SN_U8, ModelDesc, check, lib

_DT = {torch.float32: SN_F32, torch.float64: SN_F64}


def _stream() -> int:
    """raw cudaStream_t of torch's current stream on the current device (torch.cuda.current_stream() costs ~10 us of
    host time per call; seven calls per step were a fifth of the eager step)"""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


_STREAM_OBJS: dict = {}


def current_stream_obj(device) -> torch.cuda.Stream:
    """torch.cuda.current_stream(device) through a cache keyed by the raw handle (the stock call is ~10 us)"""
    idx = device.index if device.index is not None else torch._C._cuda_getDevice()
    raw = torch._C._cuda_getCurrentRawStream(idx)
    key = (idx, raw)
    st = _STREAM_OBJS.get(key)
    if st is None:
        st = torch.cuda.current_stream(device)
        if st.cuda_stream != raw:  # cannot happen; be safe
            return st
        _STREAM_OBJS[key] = st
    return st


class _on_device:
    """`with _on_device(d)` without the cost when `d` is already the current device (the common case)"""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch._C._cuda_getDevice()
        self.prev = -1

    def __enter__(self):
        cur = torch._C._cuda_getDevice()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"scenenet_b200: `{name}` must be a CUDA tensor (there is no CPU fallback); got {t.device}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------- model spec
N_PARAMS = {0: 2, 1: 2, 2: 5, 3: 5, 4: 3, 5: 3}


@dataclass
class ObserverSpec:
    """Static description of an observer: which operators, in which order, which lambda is last."""
    kinds: Sequence[int]                 # SN_KIND_* per operator, channel order
    kernel_size: Sequence[int]           # (kz, kx, ky)
    lambda_sum_order: Optional[Sequence[int]] = None  # operator ids in lambdas_dict order
    last_lambda: int = -1                # operator id; -1 = none
    observer: bool = True                # False: bare kernels (no lambdas)
    unused: frozenset = frozenset()      # param_ptrs indices that do not enter the graph (grad None, not 0)

    def n_param_ptrs(self) -> int:
        n = sum(N_PARAMS[k] for k in self.kinds)
        return n + (len(self.kinds) if self.observer else 0)

    def desc(self) -> ModelDesc:
        d = self.__dict__.get("_desc_cache")
        if d is None:
            d = self._build_desc()
            self.__dict__["_desc_cache"] = d
        return d

    def _build_desc(self) -> ModelDesc:
        G = len(self.kinds)
        if G < 1 or G > _lib.SN_MAX_GENEOS:
            raise ValueError(f"between 1 and {_lib.SN_MAX_GENEOS} GENEO operators are supported, got {G}")
        d = ModelDesc()
        d.n_geneos = G
        d.kz, d.kx, d.ky = (int(v) for v in self.kernel_size)
        off = 0
        for g, k in enumerate(self.kinds):
            d.kind[g] = k
            d.param_index[g] = off
            off += N_PARAMS[k]
        for g in range(G):
            d.lambda_index[g] = off + g if self.observer else -1
            d.lambda_sum_order[g] = (self.lambda_sum_order[g] if self.lambda_sum_order is not None else g) if self.observer else 0
        d.n_param_ptrs = off + (G if self.observer else 0)
        d.last_lambda = self.last_lambda if self.observer else -1
        return d

    @property
    def taps(self) -> int:
        kz, kx, ky = self.kernel_size
        return int(kz) * int(kx) * int(ky)


def _ptr_array(params: Sequence[torch.Tensor]):
    arr = (C.c_void_p * len(params))()
    for i, p in enumerate(params):
        if p.dtype != torch.float32 or p.numel() != 1:
            raise TypeError("GENEO parameters must be 0-dim float32 tensors")
        _need_cuda(p, "parameter")
        arr[i] = p.data_ptr()
    return arr


def _snapshot_ptr_array(snapshot: torch.Tensor):
    n = snapshot.numel()
    arr = (C.c_void_p * n)()
    base = snapshot.data_ptr()
    for i in range(n):
        arr[i] = base + 4 * i
    return arr


def unused_cone_params(kind: int, hc: int, kz: int, base: int) -> set:
    """Parameters of a cone/arrow operator that do not enter the reference's autograd graph for this
    int(apex): with hc == 0 there is no cylinder plane (arrow: `radius` unused; cone_kernel: `sigma`
    unused), with hc == kz there is no cone slice (`cone_inc`, `cone_radius` unused).  The reference
    leaves their .grad at None; so do we.  Order: apex, cone_inc, cone_radius, radius, sigma."""
    out = set()
    if kind not in (2, 3):
        return out
    if hc <= 0:
        out.add(base + (3 if kind == 3 else 4))
    if hc >= kz:
        out.update((base + 1, base + 2))
    return out


# ------------------------------------------------------------------------------- synthesis
def synth_fwd(spec: ObserverSpec, params: Sequence[torch.Tensor], write_last_lambda: bool = False,
              stream: Optional[torch.cuda.Stream] = None):
    """-> (K [G,kz,kx,ky] f32, lambda_eff [G] f32 | None, Kstar [kz,kx,ky] f32 | None, snapshot [n] f32)
    stream: launch on this stream (the caller orders it against the current one); the outputs belong to the current stream"""
    d = spec.desc()
    if len(params) != d.n_param_ptrs:
        raise ValueError(f"expected {d.n_param_ptrs} parameter tensors, got {len(params)}")
    dev = params[0].device
    G, T = len(spec.kinds), spec.taps
    # one allocation for the four small outputs (views into it): K | Kstar | lambda_eff | snapshot, each 16-byte aligned
    Tp = (T + 3) & ~3
    n_k, n_ks, n_l, n_s = G * Tp, (Tp if spec.observer else 0), (((G + 3) & ~3) if spec.observer else 0), d.n_param_ptrs
    buf = torch.empty(n_k + n_ks + n_l + n_s, dtype=torch.float32, device=dev)
    K = buf[:G * T].view(G, *spec.kernel_size)
    Kstar = buf[n_k:n_k + T].view(tuple(spec.kernel_size)) if spec.observer else None
    lam = buf[n_k + n_ks:n_k + n_ks + G] if spec.observer else None
    snap = buf[n_k + n_ks + n_l:]
    with _on_device(dev):
        check(lib.sn_geneo_synth_fwd(C.byref(d), _ptr_array(params), K.data_ptr(), _ptr(lam), _ptr(Kstar),
                                     snap.data_ptr(), int(write_last_lambda), _stream() if stream is None else stream.cuda_stream),
              "sn_geneo_synth_fwd")
    return K, lam, Kstar, snap


def synth_bwd(spec: ObserverSpec, snapshot: torch.Tensor, dK: torch.Tensor) -> torch.Tensor:
    """Jacobian^T dK -> dparams [n_param_ptrs] f32 (evaluated at the snapshot's parameter values)."""
    d = spec.desc()
    _need_cuda(dK, "dK")
    dK = dK.to(torch.float64).contiguous()
    out = torch.empty(d.n_param_ptrs, dtype=torch.float32, device=dK.device)
    with _on_device(dK.device):
        check(lib.sn_geneo_synth_bwd(C.byref(d), _snapshot_ptr_array(snapshot), dK.data_ptr(), out.data_ptr(), _stream()),
              "sn_geneo_synth_bwd")
    return out


def param_grads(spec: ObserverSpec, snapshot: torch.Tensor, K: torch.Tensor, lambda_eff: torch.Tensor, W: torch.Tensor,
                scale: float = 1.0, peer=None) -> torch.Tensor:
    """parameter gradients from the tap gradient W; `peer` (a dist.PeerAllReduce): the same launch also sums them over the
    ranks through NVLink peer memory (sn_scenenet_param_grads_allreduce)"""
    d = spec.desc()
    out = torch.empty(d.n_param_ptrs, dtype=torch.float32, device=W.device)
    with _on_device(W.device):
        if peer is None:
            check(lib.sn_scenenet_param_grads(C.byref(d), _snapshot_ptr_array(snapshot), K.data_ptr(), lambda_eff.data_ptr(),
                                              W.data_ptr(), float(scale), out.data_ptr(), _stream()), "sn_scenenet_param_grads")
        else:
            check(lib.sn_scenenet_param_grads_allreduce(C.byref(d), _snapshot_ptr_array(snapshot), K.data_ptr(), lambda_eff.data_ptr(),
                                                        W.data_ptr(), float(scale), out.data_ptr(), peer.rank, peer.world, peer.ptrs,
                                                        peer.seq.data_ptr(), peer.status.data_ptr(), peer.timeout_ms, _stream()),
                  "sn_scenenet_param_grads_allreduce")
    return out


# ------------------------------------------------------------------------------- observer
def cast_f32(x: torch.Tensor) -> torch.Tensor:
    """float64 -> float32 with our kernel; float32 passes through (made contiguous)."""
    _need_cuda(x, "x")
    if x.dtype == torch.float32:
        return x.contiguous()
    if x.dtype in (torch.uint8, torch.bool):  # occupancy grids as bytes (8x less PCIe/HBM traffic than float64)
        x = x.contiguous()
        if x.data_ptr() % 16:
            x = x.clone()
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        if x.numel():
            with _on_device(x.device):
                check(lib.sn_cast_u8_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "sn_cast_u8_to_f32")
        return out
    if x.dtype != torch.float64:
        raise TypeError(f"voxel grids must be float64, float32, uint8 or bool, got {x.dtype}")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if x.numel() == 0:
        return out
    with _on_device(x.device):
        check(lib.sn_cast_f64_to_f32(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "sn_cast_f64_to_f32")
    return out


@functools.lru_cache(maxsize=64)
def _state_words(n: int) -> int:
    return int(lib.sn_grid_state_bytes(n)) // 8


def pack_occupancy(x: torch.Tensor) -> torch.Tensor:
    """[..., Y] grid (any dtype; non-zero = occupied; Y a multiple of 32) -> int32 [..., Y / 32] with one bit per voxel
    (bit i % 32 of word i / 32 <-> flat voxel index i): the most compact form the modules accept as input — 64x fewer
    bytes than the float64 occupancy grids of the reference's ToFullDense.  Plain torch ops (not a hot path)."""
    if x.shape[-1] % 32:
        raise ValueError(f"the last dimension must be a multiple of 32 to pack rows into whole words, got {x.shape[-1]}")
    b = (x.reshape(*x.shape[:-1], x.shape[-1] // 32, 32) != 0).to(torch.int64)
    w = (b << torch.arange(32, device=x.device, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()


def unpack_occupancy(bits: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """inverse of pack_occupancy (tests, debugging)"""
    w = bits.to(torch.int64) & 0xffffffff
    return ((w.unsqueeze(-1) >> torch.arange(32, device=bits.device, dtype=torch.int64)) & 1).reshape(*bits.shape[:-1], -1).to(dtype)


def prepare(x: torch.Tensor, stream: Optional[torch.cuda.Stream] = None):
    """One HBM pass over the grid batch: -> (x32, nnz).  x32 is the float32 copy the TMA-fed stencils read
    (x itself for float32 input), nnz the grid state buffer as an int64 device tensor ([0] = number of non-zero
    voxels, [1] reserved, then one occupancy bit per voxel): the forward and the backward use
    the count ON THE DEVICE to pick the occupancy-driven kernels for sparse grids, and the occupancy-driven forward
    finds the non-zero voxels through the bits.
    stream: run the pass on this (side) stream after everything enqueued so far on the current one; the outputs
    are allocated on the current stream and the caller joins the streams (`current.wait_stream(stream)`)."""
    _need_cuda(x, "x")
    if x.dtype == torch.bool:
        x = x.view(torch.uint8)
    if x.dtype not in (torch.float64, torch.float32, torch.uint8, torch.int32):
        raise TypeError(f"voxel grids must be float64, float32, uint8, bool or int32 (packed occupancy bits), got {x.dtype}")
    x = x.contiguous()
    if x.data_ptr() % 16:
        x = x.clone()
    packed = x.dtype == torch.int32  # pack_occupancy: [..., Y / 32] words, one bit per voxel
    shape = (*x.shape[:-1], x.shape[-1] * 32) if packed else x.shape
    n = x.numel() * (32 if packed else 1)
    x32 = x if x.dtype == torch.float32 else torch.empty(shape, dtype=torch.float32, device=x.device)
    if n == 0:
        return x32, torch.zeros(8, dtype=torch.int64, device=x.device)
    nnz = torch.empty(_state_words(n), dtype=torch.int64, device=x.device)
    dt = {torch.float64: SN_F64, torch.float32: SN_F32, torch.uint8: SN_U8, torch.int32: SN_BITS}[x.dtype]
    with _on_device(x.device):
        if stream is not None:
            stream.wait_stream(current_stream_obj(x.device))  # x (and any copy made above) is ready
        st = _stream() if stream is None else stream.cuda_stream
        check(lib.sn_grid_prepare(x.data_ptr(), dt, n, None if x.dtype == torch.float32 else x32.data_ptr(),
                                  nnz.data_ptr(), st), "sn_grid_prepare")
    return x32, nnz


def _grid_dims(x: torch.Tensor):
    if x.dim() != 5 or x.shape[1] != 1:
        raise ValueError(f"expected a [B,1,Z,X,Y] voxel grid batch, got {tuple(x.shape)}")
    B, _, Z, X, Y = x.shape
    return int(B), int(Z), int(X), int(Y)


def scenenet_fwd(x32: torch.Tensor, Kstar: torch.Tensor, out_dtype: torch.dtype, nnz: Optional[torch.Tensor] = None,
                 mode: int = 0) -> torch.Tensor:
    """pred = relu(tanh(conv3d_same(x, Kstar))).  nnz (from `prepare`) enables the device-side choice of the
    occupancy-driven kernel; mode: _lib.SN_PATH_AUTO / _DENSE / _SPARSE."""
    _need_cuda(x32, "x")
    B, Z, X, Y = _grid_dims(x32)
    kz, kx, ky = (int(v) for v in Kstar.shape)
    pred = torch.empty(x32.shape, dtype=out_dtype, device=x32.device)
    if x32.numel() == 0:
        return pred
    with _on_device(x32.device):
        check(lib.sn_scenenet_fwd(x32.data_ptr(), _ptr(nnz), int(mode), Kstar.data_ptr(), B, Z, X, Y, kz, kx, ky, pred.data_ptr(),
                                  _DT[out_dtype], _stream()), "sn_scenenet_fwd")
    return pred


def scenenet_fwd_multi(x32: torch.Tensor, Kstars: torch.Tensor, out_dtype: torch.dtype, nnz: Optional[torch.Tensor] = None,
                       mode: int = 0) -> torch.Tensor:
    """several observers on the same grids (SCENENetQuantile): Kstars [Q,kz,kx,ky] float32 -> preds [Q,B,1,Z,X,Y];
    with the grid state `nnz` the occupancy-driven kernel lists the non-zero voxels once for all Q tap sets."""
    _need_cuda(x32, "x")
    B, Z, X, Y = _grid_dims(x32)
    if Kstars.dim() != 4 or Kstars.dtype != torch.float32:
        raise TypeError("Kstars must be float32 [Q, kz, kx, ky]")
    Kstars = Kstars.contiguous()
    Q, kz, kx, ky = (int(v) for v in Kstars.shape)
    preds = torch.empty((Q, *x32.shape), dtype=out_dtype, device=x32.device)
    if x32.numel() == 0:
        return preds
    with _on_device(x32.device):
        check(lib.sn_scenenet_fwd_multi(x32.data_ptr(), _ptr(nnz), int(mode), Kstars.data_ptr(), Q, B, Z, X, Y, kz, kx, ky,
                                        preds.data_ptr(), _DT[out_dtype], _stream()), "sn_scenenet_fwd_multi")
    return preds


_ws_cache: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    idx = device.index if device.index is not None else torch._C._cuda_getDevice()
    key = (idx, torch._C._cuda_getCurrentRawStream(idx))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def scenenet_bwd(x32: torch.Tensor, pred: torch.Tensor, dpred: torch.Tensor, kernel_size,
                 nnz: Optional[torch.Tensor] = None, mode: int = 0) -> torch.Tensor:
    """tap gradient W [kz,kx,ky] float64.  nnz (from `prepare`) enables the device-side choice of the
    occupancy-driven kernel."""
    B, Z, X, Y = _grid_dims(x32)
    kz, kx, ky = (int(v) for v in kernel_size)
    _need_cuda(dpred, "dpred")
    if dpred.dtype not in _DT:
        dpred = dpred.to(torch.float32)
    dpred = dpred.contiguous()
    pred = pred.contiguous()
    if dpred.data_ptr() % 16:  # vector loads in the G0 pass
        dpred = dpred.clone()
    if pred.data_ptr() % 16:
        pred = pred.clone()
    W = torch.empty((kz, kx, ky), dtype=torch.float64, device=x32.device)
    if x32.numel() == 0:
        return W.zero_()
    nbytes = int(lib.sn_scenenet_bwd_workspace_bytes(B, Z, X, Y, kz, kx, ky))
    ws = _workspace(nbytes, x32.device)
    with _on_device(x32.device):
        check(lib.sn_scenenet_bwd(x32.data_ptr(), _ptr(nnz), int(mode), pred.data_ptr(), _DT[pred.dtype], dpred.data_ptr(), _DT[dpred.dtype],
                                  B, Z, X, Y, kz, kx, ky, W.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
              "sn_scenenet_bwd")
    return W


def select_paths(x: torch.Tensor, kernel_size) -> tuple:
    """(forward mode, backward mode) the device-side selection would pick for grids like `x` — reads the grid state
    (non-zero count, clustering statistic) on the host (one synchronisation; meant for capture time, see
    graphs.GraphedStep)."""
    kz, kx, ky = (int(v) for v in kernel_size)
    x32, st = prepare(x.detach())
    B, Z, X, Y = _grid_dims(x32)
    n, dw = (int(v) for v in st[:3:2].tolist())
    return (int(lib.sn_select_fwd_path_state(n, dw, B, Z, X, Y, kz, kx, ky)), int(lib.sn_select_path(1, n, B, Z, X, Y, kz, kx, ky)))


def g0(pred: torch.Tensor, dpred: torch.Tensor) -> torch.Tensor:
    """G0 = dpred * (1 - pred^2) * [pred > 0] as float32 (pass 1 of the backward)."""
    out = torch.empty(pred.shape, dtype=torch.float32, device=pred.device)
    with _on_device(pred.device):
        check(lib.sn_scenenet_g0(pred.data_ptr(), _DT[pred.dtype], dpred.data_ptr(), _DT[dpred.dtype], pred.numel(),
                                 out.data_ptr(), _stream()), "sn_scenenet_g0")
    return out


def tapgrad(x32: torch.Tensor, g0_: torch.Tensor, kernel_size, nnz: Optional[torch.Tensor] = None, mode: int = 0) -> torch.Tensor:
    """W [kz,kx,ky] float64 from a precomputed G0 (pass 2 + 3 of the backward).
    mode: _lib.SN_TAPGRAD_AUTO / _DENSE / _SPARSE."""
    B, Z, X, Y = _grid_dims(x32)
    kz, kx, ky = (int(v) for v in kernel_size)
    W = torch.empty((kz, kx, ky), dtype=torch.float64, device=x32.device)
    ws = _workspace(int(lib.sn_scenenet_tapgrad_workspace_bytes(B, Z, X, Y, kz, kx, ky)), x32.device)
    with _on_device(x32.device):
        check(lib.sn_scenenet_tapgrad(x32.data_ptr(), g0_.data_ptr(), _ptr(nnz), int(mode), B, Z, X, Y, kz, kx, ky, W.data_ptr(), ws.data_ptr(),
                                      ws.numel(), _stream()), "sn_scenenet_tapgrad")
    return W


# ------------------------------------------------------------------------------- criterion
@dataclass
class CriterionSpec:
    """Host-side constants of the fused GENEO_Tversky_Loss (include/scenenet_b200.h: sn_criterion_fwd)."""
    ranges: Sequence[float]      # histogram bin positions (float32 values)
    w_raw: Sequence[float]       # max(1 - weight_alpha * dens_k, weight_epsilon) per bin, float32
    mse_weight: float = 1.0
    tversky_alpha: float = 0.5
    tversky_beta: float = 1.0
    focal_gamma: float = 1.0
    tversky_smooth: float = 1.0
    terms: int = 3               # bit 0: weighted MSE, bit 1: focal Tversky

    def arrays(self):
        n = len(self.ranges)
        if n != len(self.w_raw) or not 1 <= n <= _lib.SN_CRIT_MAX_BINS:
            raise ValueError(f"between 1 and {_lib.SN_CRIT_MAX_BINS} histogram bins are supported, got {n}")
        return (C.c_float * n)(*[float(v) for v in self.ranges]), (C.c_float * n)(*[float(v) for v in self.w_raw]), n


def _crit_inputs(pred: torch.Tensor, y: torch.Tensor):
    _need_cuda(pred, "y_pred")
    _need_cuda(y, "y_gt")
    if pred.dtype not in _DT:
        raise TypeError(f"criterion: float32/float64 predictions only, got {pred.dtype}")
    if y.shape != pred.shape:
        pred, y = torch.broadcast_tensors(pred, y)
    y = y.to(pred.dtype).contiguous()
    pred = pred.contiguous()
    if pred.data_ptr() % 16:
        pred = pred.clone()
    if y.data_ptr() % 16:
        y = y.clone()
    return pred, y


def criterion_fwd(pred: torch.Tensor, y: torch.Tensor, spec: CriterionSpec):
    """-> (loss [] float64, coef [SN_CRIT_COEF] float64, pred, y) — pred / y as handed to the kernel (contiguous)."""
    pred, y = _crit_inputs(pred, y)
    n = pred.numel()
    if n == 0:
        raise ValueError("criterion: empty tensors")
    loss = torch.empty((), dtype=torch.float64, device=pred.device)
    coef = torch.empty(_lib.SN_CRIT_COEF, dtype=torch.float64, device=pred.device)
    ws = _workspace(int(lib.sn_criterion_workspace_bytes(n)), pred.device)
    r, w, nb = spec.arrays()
    with _on_device(pred.device):
        check(lib.sn_criterion_fwd(pred.data_ptr(), y.data_ptr(), _DT[pred.dtype], n, r, w, nb, float(spec.mse_weight),
                                   float(spec.tversky_alpha), float(spec.tversky_beta), float(spec.focal_gamma),
                                   float(spec.tversky_smooth), int(spec.terms), loss.data_ptr(), coef.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream()), "sn_criterion_fwd")
    return loss, coef, pred, y


def criterion_bwd(pred: torch.Tensor, y: torch.Tensor, coef: torch.Tensor, spec: CriterionSpec,
                  grad_out: Optional[torch.Tensor] = None, as_g0: bool = False) -> torch.Tensor:
    """dL/dpred in pred's dtype, or (as_g0) G0 = dL/dpred * (1 - pred^2) * [pred > 0] as float32."""
    out = torch.empty(pred.shape, dtype=torch.float32 if as_g0 else pred.dtype, device=pred.device)
    if grad_out is not None:
        grad_out = grad_out.to(pred.dtype).reshape(1)
    r, w, nb = spec.arrays()
    with _on_device(pred.device):
        check(lib.sn_criterion_bwd(pred.data_ptr(), y.data_ptr(), _DT[pred.dtype], pred.numel(), r, w, nb, coef.data_ptr(),
                                   _ptr(grad_out), out.data_ptr(), int(as_g0), _stream()), "sn_criterion_bwd")
    return out


def param_penalty(params: Sequence[torch.Tensor], roles: Sequence[int], weight: float,
                  loss_accum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> float32 [2 + n]: weight * cvx_loss, weight * positive_regularizer, d(sum of both)/d param_i.
    loss_accum: float64 scalar device tensor the two penalties are ADDED to in place (the criterion's loss)."""
    if loss_accum is not None and (loss_accum.dtype != torch.float64 or loss_accum.numel() != 1 or not loss_accum.is_cuda):
        raise TypeError("loss_accum must be a float64 scalar on the device")
    n = len(params)
    if n > _lib.SN_MAX_PARAM_PTRS:
        raise ValueError(f"at most {_lib.SN_MAX_PARAM_PTRS} parameters")
    dev = params[0].device
    out = torch.empty(2 + n, dtype=torch.float32, device=dev)
    with _on_device(dev):
        check(lib.sn_param_penalty(_ptr_array(params), (C.c_int32 * n)(*[int(r) for r in roles]), n, float(weight),
                                   out.data_ptr(), _ptr(loss_accum), _stream()), "sn_param_penalty")
    return out


def threshold(p: torch.Tensor, tau: float) -> torch.Tensor:
    _need_cuda(p, "p")
    if p.dtype not in _DT:
        raise TypeError(f"threshold: float32/float64 only, got {p.dtype}")
    p = p.contiguous()
    out = torch.empty_like(p)
    with _on_device(p.device):
        check(lib.sn_threshold(p.data_ptr(), _DT[p.dtype], float(tau), p.numel(), out.data_ptr(), _stream()), "sn_threshold")
    return out


def vxg_to_xyz(vxg: torch.Tensor, origin=(0.0, 0.0, 0.0), voxel_size=(1.0, 1.0, 1.0)) -> torch.Tensor:
    """[d0*d1*d2, 4] float64 on the grid's device: (origin + index * voxel_size, value) for every voxel, C order."""
    _need_cuda(vxg, "vxg")
    if vxg.dim() != 3:
        raise ValueError(f"vxg_to_xyz: a 3-D voxel grid is expected, got shape {tuple(vxg.shape)}")
    if vxg.dtype == torch.bool:
        vxg = vxg.to(torch.uint8)
    if vxg.dtype not in (torch.float32, torch.float64, torch.uint8):
        vxg = vxg.to(torch.float64)
    vxg = vxg.contiguous()
    o = (C.c_double * 3)(*[float(v) for v in origin])
    vs = (C.c_double * 3)(*[float(v) for v in voxel_size])
    out = torch.empty((vxg.numel(), 4), dtype=torch.float64, device=vxg.device)
    if vxg.numel() == 0:
        return out
    with _on_device(vxg.device):
        check(lib.sn_vxg_to_xyz(vxg.data_ptr(), _YDT[vxg.dtype], int(vxg.shape[0]), int(vxg.shape[1]), int(vxg.shape[2]),
                                C.addressof(o), C.addressof(vs), out.data_ptr(), _stream()), "sn_vxg_to_xyz")
    return out


_YDT = {torch.float32: SN_F32, torch.float64: SN_F64, torch.uint8: SN_U8, torch.int32: SN_I32, torch.int64: SN_I64}


def confusion_counts(pred: torch.Tensor, y: torch.Tensor, tau: float, total: Optional[torch.Tensor] = None,
                     batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{TP, FP, TN, FN} of (pred >= tau) against (y != 0) in one pass (sn_confusion_counts): returns this call's counts
    as an int64 [4] device tensor (`batch`, overwritten when given) and ADDS them to `total` (int64 [4]) when given."""
    _need_cuda(pred, "pred")
    _need_cuda(y, "y")
    if pred.dtype not in _DT:
        raise TypeError(f"confusion_counts: float32/float64 predictions only, got {pred.dtype}")
    if y.dtype == torch.bool:
        y = y.view(torch.uint8)
    if y.dtype not in _YDT:
        raise TypeError(f"confusion_counts: unsupported target dtype {y.dtype}")
    if pred.numel() != y.numel():
        raise ValueError(f"confusion_counts: {pred.numel()} predictions vs {y.numel()} targets")
    pred, y = pred.contiguous(), y.contiguous()
    if pred.data_ptr() % 16:
        pred = pred.clone()
    if y.data_ptr() % 16:
        y = y.clone()
    if batch is None:
        batch = torch.empty(4, dtype=torch.int64, device=pred.device)
    for t, name in ((batch, "batch"), (total, "total")):
        if t is not None and (t.dtype != torch.int64 or t.numel() != 4 or not t.is_contiguous() or t.device != pred.device):
            raise ValueError(f"confusion_counts: `{name}` must be a contiguous int64 [4] tensor on {pred.device}")
    if pred.numel() == 0:  # nothing to count (an empty tensor has no address to hand over)
        return batch.zero_()
    with _on_device(pred.device):
        check(lib.sn_confusion_counts(pred.data_ptr(), _DT[pred.dtype], y.data_ptr(), _YDT[y.dtype], pred.numel(), float(tau),
                                      batch.data_ptr(), _ptr(total), _stream()), "sn_confusion_counts")
    return batch


def fp32_peak_probe(iters: int = 2000, device=None) -> float:
    """Measured FP32 FMA-pipe peak in TFLOP/s (roofline denominator for the stencils)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    sink = torch.zeros(1, dtype=torch.float32, device=device)
    flops = C.c_double(0.0)
    best = 0.0
    with _on_device(device):
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.sn_fp32_peak_probe(sink.data_ptr(), iters, C.byref(flops), _stream()), "sn_fp32_peak_probe")
            e1.record()
            e1.synchronize()
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
