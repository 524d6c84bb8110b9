"""FlatAdam: torch.optim.Adam semantics (no weight decay / amsgrad) as ONE kernel over flat buffers.

Parameters and gradients of the module are re-homed as views into two flat fp32 buffers (the gradient buffer is
the one :class:`myria3d_b200.parallel.FlatGradAllReducer` all-reduces), so an optimisation step is
``b200_adam_flat`` over 1.1 M contiguous floats instead of ~50 multi-tensor launches over 152 tensors.  The step
counter lives on the device: the update can be captured in a CUDA graph.  Works with
``torch.optim.lr_scheduler`` through ``param_groups[0]["lr"]`` (read at call time; re-capture after changing it
when the step is graphed)."""
from __future__ import annotations

from ctypes import c_void_p
from typing import Optional

import torch
from torch import nn

from . import _lib
from .parallel import FlatGradAllReducer


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, module: nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 reducer: Optional[FlatGradAllReducer] = None):
        params = [p for p in module.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        if not params[0].is_cuda:
            raise RuntimeError("FlatAdam runs on a CUDA (B200) device only")
        self.reducer = reducer if reducer is not None else FlatGradAllReducer(module)
        if [id(p) for p in self.reducer.params] != [id(p) for p in params]:
            raise ValueError("reducer and optimizer must cover the same parameters in the same order")
        total = self.reducer.flat.numel()
        dev = params[0].device
        self.flat_params = torch.empty(total, dtype=torch.float32, device=dev)
        offset = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.flat_params[offset:offset + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[offset:offset + n].view_as(p)  # parameters become views of the flat buffer
                offset += n
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        if not self.reducer.check_views():
            raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer "
                               "(use reducer.zero_grad() / optimizer.zero_grad(), not set_to_none=True)")
        rc = _lib.load().b200_adam_flat(
            c_void_p(self.flat_params.data_ptr()), c_void_p(self.reducer.flat.data_ptr()), c_void_p(self.exp_avg.data_ptr()),
            c_void_p(self.exp_avg_sq.data_ptr()), self.flat_params.numel(), float(g["lr"]), float(g["betas"][0]),
            float(g["betas"][1]), float(g["eps"]), c_void_p(self.step_count.data_ptr()),
            c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "b200_adam_flat")
        return loss

    def zero_grad(self, set_to_none: bool = False):  # gradients must keep aliasing the flat buffer
        self.reducer.zero_grad()
