"""CUDA-graph capture of the SCENE-Net step.

At B200 speed one fwd+bwd of a 32-grid batch is ~250 us of GPU work in 7 kernels, which is less
than what Python + autograd spend launching them: the eager module path is host-bound.  A step
on static buffers is therefore captured once (our kernels are launched on torch's current
stream, so `torch.cuda.graph` records them like any other) and replayed with one
cudaGraphLaunch.  The graph contains: [H2D copy of x] -> (kernel synthesis || grid preparation) ->
observer forward -> [criterion] -> G0 -> tap gradient -> row sum -> parameter
Jacobian -> [gradient all-reduce] -> [D2H copy of the gradients].
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedStep:
    """forward + backward of `model` on static buffers, replayable.

    Capture BEFORE running an eager backward of the same parameters on the default stream (a torch
    restriction: autograd bookkeeping created on the legacy stream cannot be joined from a capturing
    stream), or run those eager steps inside a side stream.

    x:        static device input [B,1,Z,X,Y] (float32/float64); refill it (or `x_host`) between replays
    dpred:    static upstream gradient (same shape), or None when `loss_fn(pred) -> scalar` is given
    loss_from_input: alternative to both: `x -> (loss, pred)` (e.g. `lambda x: criterion.training_loss(model, x, y)`)
    x_host:   optional pinned host tensor; when given every replay starts with x.copy_(x_host) (H2D)
    grads_host: optional pinned host float32 tensor [n_trainable]; when given every replay ends with the
              D2H copy of the flat gradient vector into it
    post_backward: optional callable run (and captured) after backward, e.g. the gradient all-reduce
    specialize: read the occupancy of `x` once before capture and enqueue only the kernels the device-side
              selection would pick for such grids (dense stencil or occupancy-driven kernel); replays on
              other data stay correct
    """

    def __init__(self, model: torch.nn.Module, x: torch.Tensor, dpred: Optional[torch.Tensor] = None,
                 loss_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, x_host: Optional[torch.Tensor] = None,
                 grads_host: Optional[torch.Tensor] = None, post_backward: Optional[Callable[[], None]] = None,
                 warmup: int = 3, specialize: bool = False,
                 loss_from_input: Optional[Callable[[torch.Tensor], tuple]] = None):
        if (dpred is not None) + (loss_fn is not None) + (loss_from_input is not None) != 1:
            raise ValueError("give exactly one of dpred / loss_fn / loss_from_input")
        self.loss_from_input = loss_from_input
        self.model, self.x, self.dpred, self.loss_fn = model, x, dpred, loss_fn
        self.x_host, self.grads_host, self.post_backward = x_host, grads_host, post_backward
        # specialize: read the occupancy of `x` once (host sync, before capture) and enqueue only the kernels the
        # device-side selection would pick for grids like it — no gated-out launches in the graph.  Both kernel
        # families are correct at any occupancy, so replaying on other data stays correct (only slower).
        self.path_modes = None
        if specialize and hasattr(model, "path_modes") and x.numel() and (x_host is None):
            from . import ops
            self.path_modes = ops.select_paths(x, model._spec_and_params()[0].kernel_size)
        self.params: Sequence[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        self.pred = None
        self.loss = None
        self.flat_grads = None
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        self.graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        from . import _lib
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._step()
        #: kernels of this library inside the captured step (the library counts its launches; a replay re-issues these)
        self.kernels_per_replay = _lib.launch_count() - n0

    def _step(self):
        if self.x_host is not None:
            self.x.copy_(self.x_host, non_blocking=True)
        saved = None
        if self.path_modes is not None:
            saved, self.model.path_modes = self.model.path_modes, self.path_modes
        try:
            if self.loss_from_input is not None:  # e.g. criterion.training_loss(model, x, y): one autograd node
                self.loss, self.pred = self.loss_from_input(self.x)
            else:
                self.pred = self.model(self.x)
        finally:
            if saved is not None:
                self.model.path_modes = saved
        # torch.autograd.grad instead of .backward(): no AccumulateGrad nodes take part, so parameters that
        # were already used eagerly (their AccumulateGrad lives on the legacy stream) cannot break the capture
        if self.loss_from_input is not None:
            grads = torch.autograd.grad(self.loss, self.params, allow_unused=True)
        elif self.loss_fn is not None:
            self.loss = self.loss_fn(self.pred)
            grads = torch.autograd.grad(self.loss, self.params, allow_unused=True)
        else:
            grads = torch.autograd.grad(self.pred, self.params, grad_outputs=self.dpred, allow_unused=True)
        self.grads = list(grads)
        for p, g in zip(self.params, grads):
            p.grad = g
        if self.post_backward is not None:
            self.post_backward()
        if self.grads_host is not None:
            self.flat_grads = torch.stack([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(()) for p in self.params])
            self.grads_host.copy_(self.flat_grads, non_blocking=True)

    def replay(self):
        self.graph.replay()
        for p, g in zip(self.params, self.grads):  # several GraphedSteps may share a model: re-bind this one's grads
            p.grad = g
        return self.pred
