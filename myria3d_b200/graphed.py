"""Whole-step CUDA graph for fixed-layout batches ("CUDA streams and graphs instead of a tracing compiler").

A RandLA-Net training step is ~230 library kernels plus a few hundred tiny torch kernels (BatchNorm
fold algebra, loss, optimizer): launched eagerly from Python the GPU idles ~1/3 of the time.  When
consecutive batches share their layout (same ``ptr``: e.g. myria3d tiles subsampled to a fixed point
budget, BASELINE configs[1]), the whole step -- gradient zeroing, forward, CrossEntropyLoss, backward and
the optimizer update -- is captured once and replayed.

Randomness: by default the decimation subsets (``decimation_indices``, pyg_randla_net.py:192-231) are drawn INSIDE
the graph by ``b200_decimation_draw`` (one kernel per level, device-side draw counter); optionally the reference's
per-cloud ``torch.randperm`` stream is drawn eagerly and handed to the graph through static index buffers.  Dropout
uses torch's graph-safe Philox state.

Training semantics are those of the eager loop: capturing a new layout (warm-up passes + capture) runs on a
snapshot of the parameters, optimizer state, BatchNorm buffers and draw counter which is restored before the first
replay, so the batch that triggers a capture is trained exactly once.  At most ``max_graphs`` layouts are kept
(least recently used evicted); a layout is captured only after it has been seen ``capture_after`` times, until then
(and for anything that cannot be captured) the step runs eagerly -- with myria3d's variable point budget most batches
are one-off layouts and simply take the eager path.  Layouts are identified by the HOST copy of ``ptr`` (a CPU
``batch.ptr`` or ``batch.ptr_host``); a device-only ``ptr`` is read back every step (one small sync).
"""
from __future__ import annotations

import os

from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops
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
        # prefetch(): device staging copies of the inputs, filled on the copy stream while the previous step runs
        self.staging: Optional[Dict[str, Tensor]] = None
        self.staged_for = None        # the batch object the staging buffers currently hold
        self.staged_event = None      # recorded on the copy stream after the host->device copies
        self.consumed_event = None    # recorded on the compute stream after staging -> static


class GraphedTrainStep:
    """``loss = step(batch)``: one optimisation step of ``model`` (a :class:`myria3d_b200.model.Model`)
    on ``batch`` (host -- ideally pinned -- or device ``Batch``), replaying a captured CUDA graph.

    ``optimizer`` must be capturable (``torch.optim.Adam(..., capturable=True)``).  ``reducer`` is the
    optional :class:`myria3d_b200.parallel.FlatGradAllReducer`; with more than one rank the NCCL
    all-reduce runs between the captured forward/backward graph and the captured optimizer graph.
    """

    def __init__(self, model, optimizer, reducer=None, warmup_steps: int = 3, decimation_rng: str = "fused",
                 max_graphs: int = 8, capture_after: int = 0):
        self.model = model
        self.max_graphs = max(1, int(max_graphs))
        self.capture_after = max(0, int(capture_after))
        self._seen: Dict[Tuple[int, ...], int] = {}
        self.net = model.model
        self.optimizer = optimizer
        self.reducer = reducer
        self.warmup_steps = warmup_steps
        # "fused" (default): one batched draw per level; "reference": the reference's per-cloud randperm stream
        self.net.decimation_rng = decimation_rng
        self.device = next(model.parameters()).device
        self._captured: "OrderedDict[Tuple[int, ...], _Captured]" = OrderedDict()
        self.library_launches = 0  # kernels of libb200randla replayed so far
        self._copy_stream: Optional["torch.cuda.Stream"] = None
        # N > 1, opt-in (B200_COLLECTIVES_IN_GRAPH=1): capture the NCCL collectives with the step -- one graph, the
        # BatchNorm-statistics broadcast hidden behind the backward.  Measured on 2 x B200: 6.43 ms/step against 6.44 for
        # the default graph | eager all-reduce | graph (the cost is the collective itself, not the host gaps), same
        # losses, and one of three runs did not shut down cleanly -- hence off by default.
        self.collectives_in_graph = os.environ.get("B200_COLLECTIVES_IN_GRAPH", "0") == "1"

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

    def _state_tensors(self) -> List[Tensor]:
        """Everything a training step mutates: parameters, optimizer state, module buffers (BatchNorm running
        statistics / counters), the fused-draw counter."""
        ts: List[Tensor] = [p.data for p in self.model.parameters()] + list(self.model.buffers())
        for attr in ("flat_params", "exp_avg", "exp_avg_sq", "step_count"):
            t = getattr(self.optimizer, attr, None)
            if isinstance(t, Tensor):
                ts.append(t)
        for st in self.optimizer.state.values():
            ts.extend(v for v in st.values() if isinstance(v, Tensor))
        if getattr(self.net, "_draw_counter", None) is not None:
            ts.append(self.net._draw_counter)
        seen, out = set(), []
        for t in ts:  # parameters of FlatAdam are views of flat_params: keep one copy per storage range
            k = (t.data_ptr(), t.numel(), t.dtype)
            if k not in seen:
                seen.add(k)
                out.append(t)
        return out

    def _capture(self, batch, key=None) -> _Captured:
        """Warm-up + capture on a snapshot of the training state (restored in place afterwards)."""
        tensors = self._state_tensors()
        if getattr(self.net, "_draw_counter", None) is None and self.net.decimation_rng == "fused":
            self.net.draw_decimation(self.net.levels_for([int(v) for v in (key or batch.ptr.tolist())], self.device), 0)
            tensors = self._state_tensors()  # the draw counter exists now
        snap = [(t, t.clone()) for t in tensors]
        known = {(t.data_ptr(), t.numel(), t.dtype) for t in tensors}
        try:
            cap = self._capture_impl(batch, key)
        finally:
            with torch.no_grad():
                for t, s0 in snap:
                    t.copy_(s0)
                # optimizer state created lazily by the warm-up steps (torch.optim.Adam): back to its initial zeros
                for t in self._state_tensors():
                    if (t.data_ptr(), t.numel(), t.dtype) not in known:
                        t.zero_()
        return cap

    def _capture_impl(self, batch, key=None) -> _Captured:
        cap = _Captured()
        dev = self.device
        cap.ptr_host = [int(v) for v in (key if key is not None else batch.ptr.tolist())]
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
                if multi and self.collectives_in_graph:
                    # ONE graph per step, collectives included: rank 0's BatchNorm statistics travel while the
                    # backward runs (they are final once the forward is done; torch DDP sends them before the NEXT
                    # forward -- same values), the flat gradient all-reduce follows the last backward kernel
                    self._zero_grad()
                    out = self.model.training_step(Batch(**cap.static), 0)
                    bcast = self.reducer.start_broadcast()
                    out["loss"].backward()
                    red = self.reducer.all_reduce(async_op=True)
                    self.reducer.finish_broadcast(bcast)
                    red.wait()
                    self.optimizer.step()
                else:
                    out = self._fwd_bwd(cap)
                    if not multi:
                        self.optimizer.step()
                cap.loss = out["loss"].detach()
                cap.logits = out["logits"].detach()
                if cap.draws_in_graph:  # the subsets of the last replay (static tensors owned by the graph)
                    cap.idx_static = list(self.net.last_decimation_idx)
            cap.launches_per_step = _lib.launch_count() - n0
            if multi and not self.collectives_in_graph:
                cap.opt_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cap.opt_graph, capture_error_mode=mode):
                    self.optimizer.step()
        finally:
            self.net.static_ptr_host, self.net.injected_decimation_idx = prev_static, prev_inj
        return cap

    @staticmethod
    def _layout_key(batch) -> Tuple[int, ...]:
        """``tuple(ptr)`` from the HOST: a CPU ``batch.ptr``, else ``batch.ptr_host``; a device-only ``ptr`` costs one
        small device->host read per step (a storage address is not an identity: the caching allocator recycles it)."""
        ptr = batch.ptr
        if not ptr.is_cuda:
            return tuple(int(v) for v in ptr.tolist())
        if "ptr_host" in batch:
            return tuple(int(v) for v in batch.ptr_host)
        return tuple(int(v) for v in ptr.tolist())

    def _eager_step(self, batch) -> Tensor:
        """The same step without a graph (one-off layouts)."""
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        self._zero_grad()
        b = batch.to(self.device, non_blocking=True) if not batch.pos.is_cuda else batch
        out = self.model.training_step(b, 0)
        out["loss"].backward()
        if self.reducer is not None:
            self.reducer.all_reduce()
        self.optimizer.step()
        self._last_eager = {"loss": out["loss"].detach(), "logits": out["logits"].detach(), "targets": b.y}
        return self._last_eager["loss"]

    # ------------------------------------------------------------------ input prefetch
    def prefetch(self, batch) -> bool:
        """Start the host->device copy of the NEXT step's (pinned) host batch on a separate stream, so that it runs under
        the step that is executing; ``step(batch)`` with the same batch object then only waits for that copy and moves
        the staged tensors into the graph's static inputs (a ~5 us device-to-device copy).  Returns False (and does
        nothing) for layouts that are not captured yet -- ``step(batch)`` copies directly then, as without prefetch.

            step.prefetch(batches[0])
            for i, b in enumerate(batches):
                loss = step(b)
                if i + 1 < len(batches):
                    step.prefetch(batches[i + 1])      # overlaps with the replay that was just enqueued
                log(loss.item())
        """
        cap = self._captured.get(self._layout_key(batch))
        if cap is None or batch.pos.is_cuda:
            return False
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        if cap.staging is None:
            cap.staging = {k: torch.empty_like(cap.static[k]) for k in ("x", "pos", "y", "batch")}
        cs = self._copy_stream
        if cap.consumed_event is not None:
            cs.wait_event(cap.consumed_event)  # the previous staged batch has been moved out
        with torch.cuda.stream(cs):
            for k in ("x", "pos", "y", "batch"):
                cap.staging[k].copy_(getattr(batch, k), non_blocking=True)
            cap.staged_event = torch.cuda.Event()
            cap.staged_event.record(cs)
        cap.staged_for = batch
        return True

    # ------------------------------------------------------------------ call
    def __call__(self, batch) -> Tensor:
        key = self._layout_key(batch)
        multi = self.reducer is not None and self.reducer.world_size > 1
        cap = self._captured.get(key)
        if multi and not (self.collectives_in_graph and cap is not None):
            self.reducer.broadcast_buffers()  # torch DDP's broadcast_buffers=True (rank 0's BatchNorm statistics)
        if cap is None:
            n_seen = self._seen.get(key, 0)
            if n_seen < self.capture_after:
                if len(self._seen) > 4096:
                    self._seen.clear()
                self._seen[key] = n_seen + 1
                self._last_key = None
                return self._eager_step(batch)
            while len(self._captured) >= self.max_graphs:  # least recently used layout goes (frees its memory pool)
                self._captured.popitem(last=False)
            cap = self._captured[key] = self._capture(batch, key)
        else:
            self._captured.move_to_end(key)
        self._last_key = key
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        if cap.staged_for is batch and cap.staged_event is not None:  # prefetch()ed: wait for the copy, move it in
            main = torch.cuda.current_stream()
            main.wait_event(cap.staged_event)
            for k in ("x", "pos", "y", "batch"):
                cap.static[k].copy_(cap.staging[k], non_blocking=True)
            cap.consumed_event = torch.cuda.Event()
            cap.consumed_event.record(main)
            cap.staged_for = None
        else:
            for k in ("x", "pos", "y", "batch"):
                cap.static[k].copy_(getattr(batch, k), non_blocking=True)
        if not cap.draws_in_graph:
            levels = self.net.levels_for(cap.ptr_host, self.device)
            idx = cap.idx_next if cap.idx_next is not None else self._draw_decimation(levels)
            for dst, src in zip(cap.idx_static, idx):
                dst.copy_(src, non_blocking=True)
        cap.graph.replay()
        ops.mark_scratch_dirty(self.device)  # the replay left its reduction sums in the scratch slices of the capture
        if cap.opt_graph is not None:
            self.reducer.all_reduce()
            cap.opt_graph.replay()
        self.library_launches += cap.launches_per_step
        if not cap.draws_in_graph:
            # draw the next step's subsets now: the launches hide behind the replay that was just enqueued
            cap.idx_next = self._draw_decimation(levels)
        return cap.loss

    def last_outputs(self, batch) -> Dict[str, Tensor]:
        cap = self._captured.get(self._layout_key(batch))
        if cap is None or getattr(self, "_last_key", None) is None:
            return dict(self._last_eager)
        return {"loss": cap.loss, "logits": cap.logits, "targets": cap.static["y"]}
