"""FlatAdam: torch.optim.Adam semantics (no weight decay / amsgrad) as ONE kernel over flat buffers.

Parameters and gradients of the module are re-homed as views into two flat fp32 buffers (the gradient buffer is
the one :class:`myria3d_b200.parallel.FlatGradAllReducer` all-reduces), so an optimisation step is
``b200_adam_flat`` over 1.1 M contiguous floats instead of ~50 multi-tensor launches over 152 tensors.  The step
counter AND the learning rate live on the device: the update can be captured in a CUDA graph and still follows a
``torch.optim.lr_scheduler`` (``sync_lr()`` copies ``param_groups[0]["lr"]`` into the device scalar; ``step()`` and
``GraphedTrainStep`` call it).  ``state_dict()`` / ``load_state_dict()`` carry the flat moments and the step counter,
so Lightning checkpoints resume with the right bias correction."""
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
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        self._lr_pinned = torch.empty(1, dtype=torch.float32).pin_memory()

    def sync_lr(self) -> None:
        """Copy ``param_groups[0]["lr"]`` to the device scalar the kernel reads (only when it changed)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr_pinned[0] = lr
            self.lr_dev.copy_(self._lr_pinned, non_blocking=True)
            self._lr_host = lr

    def check_param_views(self) -> bool:
        """True while every parameter still lives inside ``flat_params`` (``module.to()`` / ``.float()`` / a
        ``load_state_dict(assign=True)`` would re-home them and the kernel would update a buffer nobody reads)."""
        lo = self.flat_params.data_ptr()
        hi = lo + self.flat_params.numel() * 4
        return all(lo <= p.data_ptr() < hi for p in self.reducer.params)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        if not self.reducer.check_views():
            raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer "
                               "(use reducer.zero_grad() / optimizer.zero_grad(), not set_to_none=True)")
        if not self.check_param_views():
            raise RuntimeError("a parameter no longer aliases FlatAdam's flat parameter buffer (the module was moved or "
                               "cast after the optimizer was built): build the optimizer after module.to(device)")
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        rc = _lib.load().b200_adam_flat(
            c_void_p(self.flat_params.data_ptr()), c_void_p(self.reducer.flat.data_ptr()), c_void_p(self.exp_avg.data_ptr()),
            c_void_p(self.exp_avg_sq.data_ptr()), self.flat_params.numel(), float(g["lr"]), c_void_p(self.lr_dev.data_ptr()),
            float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), c_void_p(self.step_count.data_ptr()),
            c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "b200_adam_flat")
        return loss

    # ---- checkpointing: torch.optim.Optimizer.state_dict() only knows `self.state`, which FlatAdam does not use
    def state_dict(self):
        sd = super().state_dict()
        sd["flat_adam"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                           "step": self.step_count.detach().clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("flat_adam", None)
        super().load_state_dict(state_dict)
        if flat is not None:
            if flat["exp_avg"].numel() != self.exp_avg.numel():
                raise ValueError("FlatAdam.load_state_dict: moment buffers of a different parameter count")
            self.exp_avg.copy_(flat["exp_avg"])
            self.exp_avg_sq.copy_(flat["exp_avg_sq"])
            self.step_count.copy_(flat["step"])
        self._lr_host = float("nan")  # force the next sync_lr()

    def zero_grad(self, set_to_none: bool = False):  # gradients must keep aliasing the flat buffer
        self.reducer.zero_grad()
