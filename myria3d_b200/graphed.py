"""Whole-step CUDA graph for fixed-layout batches ("CUDA streams and graphs instead of a tracing compiler").

A RandLA-Net training step is ~230 library kernels plus a few hundred tiny torch kernels (BatchNorm
fold algebra, loss, optimizer): launched eagerly from Python the GPU idles ~1/3 of the time.  When
consecutive batches share their layout (same ``ptr``: e.g. myria3d tiles subsampled to a fixed point
budget, BASELINE configs[1]), the whole step -- gradient zeroing, forward, CrossEntropyLoss, backward and
the optimizer update -- is captured once and replayed.

Randomness stays outside the capture: the decimation subsets (``decimation_indices``,
pyg_randla_net.py:192-231) are drawn eagerly -- by default with one batched draw per level
(``fused_decimation_indices``), optionally with the reference's per-cloud ``torch.randperm`` stream -- and
handed to the graph through static index buffers; dropout uses torch's graph-safe Philox state.  One graph is
kept per batch layout.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .data import Batch


class _Captured:
    def __init__(self):
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.opt_graph: Optional[torch.cuda.CUDAGraph] = None
        self.static: Dict[str, Tensor] = {}
        self.idx_static: List[Tensor] = []
        self.idx_next: Optional[List[Tensor]] = None
        self.draws_in_graph = False
        self.loss: Optional[Tensor] = None
        self.logits: Optional[Tensor] = None
        self.ptr_host: List[int] = []
        self.launches_per_step = 0


class GraphedTrainStep:
    """``loss = step(batch)``: one optimisation step of ``model`` (a :class:`myria3d_b200.model.Model`)
    on ``batch`` (host -- ideally pinned -- or device ``Batch``), replaying a captured CUDA graph.

    ``optimizer`` must be capturable (``torch.optim.Adam(..., capturable=True)``).  ``reducer`` is the
    optional :class:`myria3d_b200.parallel.FlatGradAllReducer`; with more than one rank the NCCL
    all-reduce runs between the captured forward/backward graph and the captured optimizer graph.
    """

    def __init__(self, model, optimizer, reducer=None, warmup_steps: int = 3, decimation_rng: str = "fused"):
        self.model = model
        self.net = model.model
        self.optimizer = optimizer
        self.reducer = reducer
        self.warmup_steps = warmup_steps
        # "fused" (default): one batched draw per level; "reference": the reference's per-cloud randperm stream
        self.net.decimation_rng = decimation_rng
        self.device = next(model.parameters()).device
        self._captured: Dict[Tuple[int, ...], _Captured] = {}
        self._key_cache: Dict[Tuple[int, int], Tuple[int, ...]] = {}
        self.library_launches = 0  # kernels of libb200randla replayed so far

    # ------------------------------------------------------------------ helpers
    def _zero_grad(self):
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=False)

    def _draw_decimation(self, levels) -> List[Tensor]:
        """The four levels' random subsets, level-major (same order of draws as the reference's forward)."""
        return [self.net.draw_decimation(levels, l) for l in range(4)]

    def _fwd_bwd(self, cap: _Captured):
        self._zero_grad()
        out = self.model.training_step(Batch(**cap.static), 0)
        out["loss"].backward()
        return out

    def _capture(self, batch) -> _Captured:
        cap = _Captured()
        dev = self.device
        cap.ptr_host = [int(v) for v in batch.ptr.tolist()]
        for k in ("x", "pos", "y", "batch", "ptr"):
            cap.static[k] = getattr(batch, k).to(dev, non_blocking=True).clone()
        levels = self.net.levels_for(cap.ptr_host, dev)
        # "fused" draws (randint + sort + gather per level) are graph-safe: they are captured with the step, every
        # replay advances the Philox offset and draws fresh subsets -- no eager launches between replays.  The
        # reference's per-cloud randperm stream stays outside the graph and enters through static index buffers.
        cap.draws_in_graph = self.net.decimation_rng == "fused"
        cap.idx_static = None if cap.draws_in_graph else self._draw_decimation(levels)
        multi = self.reducer is not None and self.reducer.world_size > 1

        prev_static, prev_inj = self.net.static_ptr_host, self.net.injected_decimation_idx
        self.net.static_ptr_host, self.net.injected_decimation_idx = cap.ptr_host, cap.idx_static
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup_steps):
                    self._fwd_bwd(cap)
                    if multi:
                        self.reducer.all_reduce()
                    self.optimizer.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()

            n0 = _lib.launch_count()
            # with a process group alive, NCCL's watchdog thread may touch the CUDA API while we capture: only calls of
            # THIS thread may invalidate the capture then
            mode = "thread_local" if multi else "global"
            cap.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cap.graph, capture_error_mode=mode):
                out = self._fwd_bwd(cap)
                if not multi:
                    self.optimizer.step()
                cap.loss = out["loss"].detach()
                cap.logits = out["logits"].detach()
                if cap.draws_in_graph:  # the subsets of the last replay (static tensors owned by the graph)
                    cap.idx_static = list(self.net.last_decimation_idx)
            cap.launches_per_step = _lib.launch_count() - n0
            if multi:
                cap.opt_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cap.opt_graph, capture_error_mode=mode):
                    self.optimizer.step()
        finally:
            self.net.static_ptr_host, self.net.injected_decimation_idx = prev_static, prev_inj
        return cap

    def _layout_key(self, ptr: Tensor) -> Tuple[int, ...]:
        """``tuple(ptr)``; device tensors are read back once per (storage, version) to avoid a sync per step."""
        if not ptr.is_cuda:
            return tuple(int(v) for v in ptr.tolist())
        ck = (ptr.data_ptr(), ptr._version)
        key = self._key_cache.get(ck)
        if key is None:
            if len(self._key_cache) > 64:
                self._key_cache.clear()
            key = self._key_cache[ck] = tuple(int(v) for v in ptr.tolist())
        return key

    # ------------------------------------------------------------------ call
    def __call__(self, batch) -> Tensor:
        key = self._layout_key(batch.ptr)
        cap = self._captured.get(key)
        if cap is None:
            cap = self._captured[key] = self._capture(batch)
        for k in ("x", "pos", "y", "batch"):
            cap.static[k].copy_(getattr(batch, k), non_blocking=True)
        if not cap.draws_in_graph:
            levels = self.net.levels_for(cap.ptr_host, self.device)
            idx = cap.idx_next if cap.idx_next is not None else self._draw_decimation(levels)
            for dst, src in zip(cap.idx_static, idx):
                dst.copy_(src, non_blocking=True)
        cap.graph.replay()
        if cap.opt_graph is not None:
            self.reducer.all_reduce()
            cap.opt_graph.replay()
        self.library_launches += cap.launches_per_step
        if not cap.draws_in_graph:
            # draw the next step's subsets now: the launches hide behind the replay that was just enqueued
            cap.idx_next = self._draw_decimation(levels)
        return cap.loss

    def last_outputs(self, batch) -> Dict[str, Tensor]:
        cap = self._captured[self._layout_key(batch.ptr)]
        return {"loss": cap.loss, "logits": cap.logits, "targets": cap.static["y"]}
