"""Base class of the parametric GENEO kernels — host-side mirror of the reference's
core/models/geneos/GENEO_kernel_torch.py:17-116 (same constructor contract: building the
object computes `.kernel`), with the arithmetic done by one CUDA launch
(csrc/synth.cu::synth_fwd_kernel) instead of ~30-60 tiny ATen ops, and a closed-form
Jacobian^T (synth_bwd_kernel) instead of autograd through them.
"""
from __future__ import annotations

from abc import abstractmethod

import torch

from .... import ops
from ...._lib import KIND


class _KernelSynthesis(torch.autograd.Function):
    """kernel = synth(kind, kernel_size, *params); differentiable w.r.t. the float params."""

    @staticmethod
    def forward(ctx, spec, *params):
        K, _, _, snap = ops.synth_fwd(spec, [p.detach() for p in params])
        ctx.spec = spec
        ctx.save_for_backward(snap)
        return K[0]

    @staticmethod
    def backward(ctx, dK):
        (snap,) = ctx.saved_tensors
        d = ops.synth_bwd(ctx.spec, snap, dK.reshape(1, -1))
        return (None, *[d[i] if (ctx.needs_input_grad[i + 1] and i not in ctx.spec.unused) else None
                        for i in range(d.numel())])


def _as_param(v, device):
    if not torch.is_tensor(v):
        v = torch.tensor(float(v))
    return v.to(device=device, dtype=torch.float32).reshape(())


class GENEO_kernel_torch:
    """
    Initialization class for GENEO kernels (3-D arrays convolved with voxel grids).

    * kernel shape in (z, x, y)
    """

    #: name of the kernel family in the C ABI (SN_KIND_*), set by subclasses
    kind_name: str = ""
    #: alphabetical parameter order = the order of the C ABI (nn.ParameterDict order)
    abi_params: tuple = ()

    def __init__(self, name, kernel_size, plot=False):
        self.name = name
        self.kernel_size = kernel_size
        self.plot = plot
        if not torch.cuda.is_available():
            raise RuntimeError("scenenet_b200 GENEO kernels are synthesised on the GPU; no CUDA device is available "
                               "and there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.volume = int(kernel_size[0]) * int(kernel_size[1]) * int(kernel_size[2])
        self.kernel = self.compute_kernel()
        if plot:
            self.plot_kernel()

    # -- values of the ABI parameters, filled by subclasses' __init__ ----------------------
    def _abi_values(self):
        return [getattr(self, p) for p in self.abi_params]

    def compute_kernel(self) -> torch.Tensor:
        """Returns the 3-D GENEO kernel [kz,kx,ky], float32, differentiable."""
        kind = KIND[self.kind_name]
        unused = frozenset()
        if hasattr(self, "apex"):
            unused = frozenset(ops.unused_cone_params(kind, int(float(self.apex)), int(self.kernel_size[0]), 0))
        spec = ops.ObserverSpec(kinds=[kind], kernel_size=tuple(int(k) for k in self.kernel_size), observer=False,
                                unused=unused)
        params = [_as_param(v, self.device) for v in self._abi_values()]
        return _KernelSynthesis.apply(spec, *params)

    def convolution(self, tensor: torch.Tensor, plot=True) -> torch.Tensor:
        """Cross-correlates the kernel with `tensor` [B,1,Z,X,Y] ('same' zero padding)."""
        # the reference's helper is a timing/plotting utility (GENEO_kernel_torch.py:46-62), not on the
        # hot path; the C ABI only exposes the fused observer forward.
        raise NotImplementedError("GENEO_kernel_torch.convolution is a plotting helper of the reference and is "
                                  "outside the hot path; use SceneNet.forward")

    def plot_kernel(self):
        print(f"\n{'*' * 50}")
        print(f"kernel shape = {tuple(self.kernel.shape)}")
        print(f"kernel sum = {torch.sum(self.kernel)}")

    @staticmethod
    def mandatory_parameters():
        return []

    @staticmethod
    def geneo_parameters():
        return []

    @staticmethod
    def geneo_smart_config():
        return

    @staticmethod
    def geneo_random_config(name='GENEO_rand'):
        """Base of every random configuration (GENEO_kernel_torch.py:97-116): kernel (9,9,9) gives the
        ranges; consumes no random numbers (the base class has no parameters)."""
        return {'name': name, 'kernel_size': (9, 9, 9), 'plot': False, 'geneo_params': {}, 'non_trainable': []}
