"""Data-parallel plumbing for the hot path: one process per GPU, tiles sharded over ranks,
ONE NCCL all-reduce of a flat fp32 gradient buffer per optimizer step (BASELINE north_star:
"NCCL allreduce on gradients only").

The reference does this through Lightning's DDP strategy
(``configs/experiment/RandLaNet_base_run_FR-MultiGPU.yaml:9-13``): 1 113 719 fp32 gradients =
4.45 MB in 1-2 buckets, BatchNorm statistics local to each rank (no SyncBN).  Here every
``param.grad`` is a view into one contiguous buffer, so the whole gradient is reduced with a single
latency-bound collective (NVLS / NVSwitch on the 8xB200 box) enqueued right after the last backward
kernel; tiles never cross ranks, so there is no other collective on the data path.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import nn


class FlatGradAllReducer:
    """Owns a flat gradient buffer for ``module`` and averages it across ranks.

    Usage::

        reducer = FlatGradAllReducer(net)          # after net.to(device)
        loss.backward()                            # grads accumulate in the flat buffer
        reducer.all_reduce()                       # one collective, async on `stream` if given
        optimizer.step(); reducer.zero_grad()
    """

    def __init__(self, module: nn.Module, process_group: Optional[dist.ProcessGroup] = None):
        self.params: List[nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameter")
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        self.group = process_group
        offset = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[offset:offset + n].view_as(p)
            offset += n
        # the backward kernels may accumulate straight into these views (ops._direct_grad): this class, not
        # autograd hooks, is what reduces them
        from . import ops

        ops.enable_direct_grads(self.flat)
        self._buffers: Optional[List[torch.Tensor]] = None
        self._flat_buffers: Optional[torch.Tensor] = None
        self._module = module

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def check_views(self) -> bool:
        """True while every ``param.grad`` still aliases the flat buffer (a ``zero_grad(set_to_none=True)``
        or an optimizer that replaces ``.grad`` would break the aliasing)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def all_reduce(self, async_op: bool = False):
        """Average the gradients over the ranks (sum, then divide by the world size)."""
        ws = self.world_size
        if ws == 1:
            return None
        self.flat.div_(ws)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def zero_grad(self) -> None:
        self.flat.zero_()

    def _buffer_views(self):
        if self._buffers is None:
            self._buffers = [b for b in self._module.buffers() if b.is_floating_point()]
            self._flat_buffers = torch.empty(sum(b.numel() for b in self._buffers), dtype=torch.float32,
                                             device=self.flat.device)
        return ([b.reshape(-1) for b in self._buffers],
                list(self._flat_buffers.split([b.numel() for b in self._buffers])))

    def broadcast_buffers(self, src: int = 0) -> None:
        """torch DDP's default ``broadcast_buffers=True``: before every forward, rank ``src``'s floating buffers (the
        BatchNorm running statistics: 8 179 floats) replace everybody else's, so that every rank checkpoints the
        same state (the reference trains with Lightning's stock DDP,
        configs/experiment/RandLaNet_base_run_FR-MultiGPU.yaml:9-13).  One flat broadcast per call; a no-op on one
        rank.  ``num_batches_tracked`` counters advance identically on every rank and are left alone."""
        self.finish_broadcast(self.start_broadcast(src))

    def start_broadcast(self, src: int = 0):
        """First half of :meth:`broadcast_buffers`: pack + asynchronous broadcast.  Returns a handle for
        :meth:`finish_broadcast` (None on one rank).  GraphedTrainStep issues it right after the forward pass -- the
        running statistics do not change again until the next forward -- so the transfer hides behind the backward."""
        if self.world_size == 1:
            return None
        bufs, parts = self._buffer_views()
        if not bufs:
            return None
        torch._foreach_copy_(parts, bufs)
        return dist.broadcast(self._flat_buffers, src=src, group=self.group, async_op=True)

    def finish_broadcast(self, work) -> None:
        if work is None:
            return
        work.wait()
        bufs, parts = self._buffer_views()
        torch._foreach_copy_(bufs, parts)


def shard_tiles(num_tiles: int, rank: int, world_size: int) -> List[int]:
    """Tile indices of ``rank`` under a DistributedSampler-like round-robin split (no shuffling,
    padded by wrapping so every rank gets the same count, like ``DistributedSampler(drop_last=False)``)."""
    if num_tiles <= 0:
        return []
    per_rank = -(-num_tiles // world_size)
    return [(rank + i * world_size) % num_tiles for i in range(per_rank)]


def broadcast_module_state(module: nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make parameters and buffers identical on every rank at start-up (what DDP's constructor does)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
