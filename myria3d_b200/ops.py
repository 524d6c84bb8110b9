"""Host-side operators of the B200 RandLA-Net hot path: thin wrappers + ``torch.autograd.Function``s
over the C ABI of ``libb200randla.so``.  PyTorch only owns device memory, streams and the autograd
graph; every computation on ``[N, .]`` / ``[E, .]`` data is a kernel of the library.

Each operator cites the reference call it stands for (paths relative to the myria3d repository).
"""
from __future__ import annotations

from ctypes import c_void_p
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.autograd.function import once_differentiable

from . import _lib

LRELU_SLOPE = 0.2  # myria3d/models/modules/pyg_randla_net.py:92
BN_MOMENTUM = 0.01  # pyg_randla_net.py:94
BN_EPS = 1e-6  # pyg_randla_net.py:94


def _p(t: Optional[Tensor]) -> Optional[c_void_p]:
    return None if t is None else c_void_p(t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _call(name: str, *args) -> None:
    """Enqueue one library entry point on the current stream (optionally timed by ``_lib.PROFILER``)."""
    lib = _lib.load()
    prof = _lib.PROFILER
    if prof is None:
        _lib.check(getattr(lib, name)(*args), name)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    rc = getattr(lib, name)(*args)
    end.record()
    _lib.check(rc, name)
    prof.add(name, args, start, end)


# Storages (untyped_storage().data_ptr()) whose owners asked for direct gradient accumulation: the flat gradient
# buffers of FlatGradAllReducer (and so of FlatAdam / GraphedTrainStep, which are built on it).
_DIRECT_GRAD_STORAGES: set = set()

# Fork/join of independent backward kernels (weight gradient || input gradient of one Linear) onto a second stream.
# Only while a CUDA graph is being captured: the fork then becomes two parallel branches of the graph (the kernels of
# levels 2-4 fill a fraction of the 148 SMs each and run side by side); in eager mode the event traffic would only
# add host time.
PARALLEL_BACKWARD = True
_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return s


def _fork_backward() -> bool:
    return PARALLEL_BACKWARD and torch.cuda.is_current_stream_capturing()


def fork_enabled() -> bool:
    """True while a CUDA graph is being captured (and PARALLEL_BACKWARD is on): independent kernel chains may then be
    enqueued on :func:`side_stream` between ``side.wait_stream(main)`` and ``main.wait_stream(side)`` / events."""
    return _fork_backward()


def side_stream(device: torch.device) -> "torch.cuda.Stream":
    return _side_stream(device)


def enable_direct_grads(flat: Tensor) -> None:
    """Allow the backward kernels to accumulate parameter gradients straight into views of ``flat``."""
    _DIRECT_GRAD_STORAGES.add(flat.untyped_storage().data_ptr())


def disable_direct_grads(flat: Tensor) -> None:
    _DIRECT_GRAD_STORAGES.discard(flat.untyped_storage().data_ptr())


def _direct_grad(p: Optional[Tensor]) -> Optional[Tensor]:
    """``p.grad`` when a backward kernel may accumulate straight into it, else None (the gradient then goes back
    through autograd's AccumulateGrad node like any other op's).

    Direct accumulation is OPT-IN: only gradients that are views of a buffer registered with
    :func:`enable_direct_grads` qualify -- the flat buffer of :class:`myria3d_b200.parallel.FlatGradAllReducer`, whose
    owner reduces it itself.  Anything that relies on AccumulateGrad running (torch DDP's reducer hooks it, so do
    ``register_post_accumulate_grad_hook`` users) therefore keeps working: with plain ``p.grad`` tensors this function
    returns None.  Saves the zero-fill of a temporary and the AccumulateGrad add per parameter; the result
    (p.grad += g) is the same."""
    if not _DIRECT_GRAD_STORAGES or p is None or not isinstance(p, torch.nn.Parameter) or not p.requires_grad:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape or not g.is_cuda:
        return None
    if g.untyped_storage().data_ptr() not in _DIRECT_GRAD_STORAGES:
        return None
    if p._backward_hooks or getattr(p, "_post_accumulate_grad_hooks", None):
        return None
    return g


class _ScratchArena:
    """Zero-filled scratch for the small reduction buffers of the backward pass (BatchNorm sums, encoder gradients, loss
    accumulators): slices of one buffer that ``reset()`` clears with ONE memset at the start of the next forward instead
    of one fill kernel per buffer (~90 launches per training step).  A slice may only be handed to consumers that finish
    within the same forward/backward pass; when the arena is full callers fall back to ``torch.zeros``.

    ``reset()`` clears everything that was EVER handed out (``hw``, the high-water mark), not just the last pass: the
    replay of a captured step graph leaves its sums in the slices it was captured with, which may lie beyond the extent
    of the pass that ran last from Python.  Replays are invisible from here, so ``GraphedTrainStep`` reports them
    (``mark_scratch_dirty``) and the next eager ``take()`` clears the arena first -- callers that never go through
    ``B200RandLANet.forward`` (a bare LocalFeatureAggregation, ``ops.cross_entropy``) stay correct after a replay."""

    def __init__(self, device: torch.device, nbytes: int = 1 << 20):
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.off = 0
        self.hw = 0
        self.dirty = False

    def take(self, numel: int, dtype: torch.dtype) -> Optional[Tensor]:
        if self.dirty and not self._capturing():
            self.reset()
        nbytes = (numel * torch.empty((), dtype=dtype).element_size() + 15) // 16 * 16
        if self.off + nbytes > self.buf.numel():
            return None
        t = self.buf[self.off:self.off + nbytes].view(dtype)[:numel]
        self.off += nbytes
        self.hw = max(self.hw, self.off)
        return t

    def reset(self) -> None:
        if self.hw:
            self.buf[:self.hw].zero_()
        self.off = 0
        if not self._capturing():
            self.dirty = False

    def _capturing(self) -> bool:
        return self.buf.is_cuda and torch.cuda.is_current_stream_capturing()


_ARENAS: dict = {}


def _arena_key(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


def reset_scratch(device: torch.device) -> None:
    """Called by ``B200RandLANet.forward``: everything handed out since the last call is dead by now."""
    a = _ARENAS.get(_arena_key(device))
    if a is not None:
        a.reset()


def mark_scratch_dirty(device: torch.device) -> None:
    """Called after the replay of a captured step graph: its kernels wrote into the arena slices of the capture."""
    a = _ARENAS.get(_arena_key(device))
    if a is not None:
        a.dirty = True


def _zeros_scratch(numel: int, dtype: torch.dtype, device: torch.device) -> Tensor:
    key = _arena_key(device)
    a = _ARENAS.get(key)
    if a is None:
        a = _ARENAS[key] = _ScratchArena(device)
    t = a.take(numel, dtype)
    return t if t is not None else torch.zeros(numel, dtype=dtype, device=device)


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "myria3d_b200 operators run on a B200 only (got a CPU tensor); there is no CPU fallback"
            )


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def table_width(k: int) -> int:
    """Neighbour-table width used by the fused LFA kernels for ``k`` neighbours."""
    if k <= 16:
        return 16
    if k <= 32:
        return 32
    raise ValueError(f"num_neighbors={k} > 32 is not supported by the fused LFA kernels")


# ------------------------------------------------------------------------------------------- kNN
GRID_KNN_MIN_POINTS = 512  # kNN graph (kt = 16/32): measured crossover ~400 points per cloud (grid 60 us vs brute 92 us at 800)


def knn(pos_x: Tensor, ptr_x: Tensor, pos_y: Tensor, ptr_y: Tensor, k: int, max_queries_per_cloud: int,
        kt: Optional[int] = None, want_dist: bool = True, max_points_per_cloud: Optional[int] = None,
        algo: str = "auto") -> Tuple[Tensor, Optional[Tensor]]:
    """k nearest ``pos_x`` points of the same cloud for every ``pos_y`` point.

    Stands for torch_cluster ``knn`` behind ``knn_graph`` (pyg_randla_net.py:180) and
    ``knn_interpolate`` (pyg_randla_net.py:250, models/model.py:90).  Returns ``nbr`` int32
    ``[ny, kt]`` (-1 padded) and ``dist2`` fp32 ``[ny, kt]`` (inf padded).

    ``algo``: ``"brute"`` = tiled brute force (TMA-staged candidate tiles), ``"grid"`` = bucket-grid ring
    search, ``"auto"`` = grid when the largest candidate cloud has >= ``GRID_KNN_MIN_POINTS`` points.
    Both are exact and return identical tables.  ``max_points_per_cloud`` (largest ``pos_x`` cloud)
    defaults to ``max_queries_per_cloud`` (right for self-queries).
    """
    _need_cuda(pos_x, ptr_x, pos_y, ptr_y)
    kt = k if kt is None else kt
    same = pos_x is pos_y and ptr_x is ptr_y
    pos_x = _f32c(pos_x)
    pos_y = pos_x if same else _f32c(pos_y)
    ptr_x = ptr_x.to(torch.int64).contiguous()
    ptr_y = ptr_x if same else ptr_y.to(torch.int64).contiguous()
    max_x = int(max_queries_per_cloud if max_points_per_cloud is None else max_points_per_cloud)
    ny, nx = pos_y.shape[0], pos_x.shape[0]
    num_clouds = ptr_x.numel() - 1
    nbr = torch.empty((ny, kt), dtype=torch.int32, device=pos_y.device)
    dist2 = torch.empty((ny, kt), dtype=torch.float32, device=pos_y.device) if want_dist else None
    if algo == "auto":
        # the warp-cooperative grid search serves neighbour tables (kt = 16 / 32); the thread-per-query grid kernel
        # behind k = 1 interpolation queries only pays off on larger clouds
        algo = "grid" if max_x >= (GRID_KNN_MIN_POINTS if kt in (16, 32) and k > 1 else 1024) else "brute"
    if algo == "grid":
        nbytes = int(_lib.load().b200_knn_grid_workspace_bytes(nx, num_clouds, max_x))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=pos_x.device)
        _call("b200_knn_grid", _p(pos_x), _p(ptr_x), nx, _p(pos_y), _p(ptr_y), ny, num_clouds, max_x,
              int(max_queries_per_cloud), k, kt, _p(nbr), _p(dist2), _p(ws), nbytes, _stream())
    elif algo == "brute":
        _call("b200_knn", _p(pos_x), _p(ptr_x), nx, _p(pos_y), _p(ptr_y), ny, num_clouds,
              int(max_queries_per_cloud), k, kt, _p(nbr), _p(dist2), _stream())
    else:
        raise ValueError(f"unknown kNN algorithm {algo!r}")
    return nbr, dist2


def edge_moments(pos: Tensor, nbr: Tensor) -> Tensor:
    """fp64 ``[57]``: edge count, sum q, sum q q^T for q = (p_i, p_j, |p_j - p_i|) over all edges."""
    _need_cuda(pos, nbr)
    out = torch.zeros(57, dtype=torch.float64, device=pos.device)
    _call("b200_edge_moments", _p(pos), _p(nbr), pos.shape[0], nbr.shape[1], _p(out), _stream())
    return out


# ------------------------------------------------------------------ encoder fold (Linear + BatchNorm -> affine)
class _EncoderFold(torch.autograd.Function):
    """mlp_encoder = Linear(10->h) + BatchNorm1d over all E edges (pyg_randla_net.py:117,144) folded into the
    affine map of q = (p_i, p_j, dist) that the fused LFA kernels consume; exact train-mode gradients."""

    @staticmethod
    def forward(ctx, w, b, gamma, beta, moments, rm, rv, nbt, momentum, eps):
        params = (w, b, gamma, beta)
        w, gamma, beta = _f32c(w), _f32c(gamma), _f32c(beta)
        b = _f32c(b) if b is not None else None
        h = w.shape[0]
        enc_w = torch.empty((h, 7), dtype=torch.float32, device=w.device)
        enc_b = torch.empty(h, dtype=torch.float32, device=w.device)
        _call("b200_encoder_fold_fwd", _p(w), _p(b), _p(gamma), _p(beta), _p(moments), _p(rm), _p(rv),
              _p(nbt) if moments is not None else None, momentum, eps, _p(enc_w), _p(enc_b), h, _stream())
        training = moments is not None
        # eval-mode backward reads the (then unchanging) running statistics; train mode reads the moments
        ctx.save_for_backward(w, b, gamma, moments, None if training else rm, None if training else rv)
        ctx.eps = eps
        ctx.params = params
        return enc_w, enc_b

    @staticmethod
    @once_differentiable
    def backward(ctx, g_enc_w, g_enc_b):
        w, b, gamma, moments, rm, rv = ctx.saved_tensors
        h = w.shape[0]
        direct = [_direct_grad(p) for p in ctx.params]
        if all(d is not None or p is None for d, p in zip(direct, ctx.params)):  # accumulate into the .grad buffers
            gw, gb, gg, gbeta = direct
            _call("b200_encoder_fold_bwd", _p(w), _p(b), _p(gamma), _p(moments), _p(rm), _p(rv), ctx.eps,
                  _p(_f32c(g_enc_w)), _p(_f32c(g_enc_b)), _p(gw), _p(gb), _p(gg), _p(gbeta), h, 1, _stream())
            return (None,) * 10
        gw = torch.empty_like(w)
        gb = torch.empty_like(b) if b is not None else None
        gg = torch.empty_like(gamma)
        gbeta = torch.empty_like(gamma)
        _call("b200_encoder_fold_bwd", _p(w), _p(b), _p(gamma), _p(moments), _p(rm), _p(rv), ctx.eps,
              _p(_f32c(g_enc_w)), _p(_f32c(g_enc_b)), _p(gw), _p(gb), _p(gg), _p(gbeta), h, 0, _stream())
        return gw, gb, gg, gbeta, None, None, None, None, None, None


def encoder_fold(lin: torch.nn.Linear, bn: torch.nn.BatchNorm1d, moments: Optional[Tensor], num_edges: int,
                 training: bool) -> Tuple[Tensor, Tensor]:
    """(enc_w [h,7], enc_b [h]) of an LFA's mlp_encoder; see ``b200_encoder_fold_fwd`` in the C header."""
    _need_cuda(lin.weight)
    if training and num_edges <= 1:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size [{num_edges}, 10]")
    return _EncoderFold.apply(lin.weight, lin.bias, bn.weight, bn.bias, moments if training else None,
                              bn.running_mean, bn.running_var, bn.num_batches_tracked, float(bn.momentum),
                              float(bn.eps))


# ------------------------------------------------------------------ LocSE + attentive pooling
class _LFAFunction(torch.autograd.Function):
    """propagate()+message() of LocalFeatureAggregation (pyg_randla_net.py:121-152), fused."""

    @staticmethod
    def forward(ctx, x, pos, nbr, enc_w, enc_b, att_w):
        ctx.att_param = att_w
        x, enc_w, enc_b, att_w = _f32c(x), _f32c(enc_w), _f32c(enc_b), _f32c(att_w)
        att_wt = att_w.t().contiguous()
        n, h = x.shape
        c = 2 * h
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        _call("b200_lfa_fwd", _p(x), _p(pos), _p(nbr), _p(enc_w), _p(enc_b), _p(att_wt), _p(out),
                                      n, c, nbr.shape[1], _stream())
        ctx.save_for_backward(x, pos, nbr, enc_w, enc_b, att_wt, att_w)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, pos, nbr, enc_w, enc_b, att_wt, att_w = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        n, h = x.shape
        c = 2 * h
        gx = torch.zeros_like(x)
        small = _zeros_scratch(enc_w.numel() + enc_b.numel(), torch.float32, x.device)  # consumed by _EncoderFold.backward
        gew = small[:enc_w.numel()].view_as(enc_w)
        geb = small[enc_w.numel():]
        direct = _direct_grad(ctx.att_param)
        gaw = direct if direct is not None else torch.zeros_like(att_w)
        nbytes = int(_lib.load().b200_lfa_bwd_workspace_bytes(n, c, nbr.shape[1]))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device) if nbytes else None
        _call("b200_lfa_bwd", _p(x), _p(pos), _p(nbr), _p(enc_w), _p(enc_b), _p(att_wt), _p(att_w), _p(grad_out),
              _p(gx), _p(gew), _p(geb), _p(gaw), _p(ws), nbytes, n, c, nbr.shape[1], _stream())
        return gx, None, None, gew, geb, (None if direct is not None else gaw)


def lfa_attentive_pool(x: Tensor, pos: Tensor, nbr: Tensor, enc_w: Tensor, enc_b: Tensor, att_w: Tensor) -> Tensor:
    """``sum_j softmax_j(W_att f_ij) * f_ij`` with ``f_ij = [x_j ; lrelu(enc_w q_ij + enc_b)]``.

    ``enc_w`` ``[c/2, 7]`` / ``enc_b`` act on ``q = (p_i, p_j, |p_j-p_i|)`` and already contain the
    encoder BatchNorm (see :func:`myria3d_b200.randla_net.fold_encoder`)."""
    _need_cuda(x, pos, nbr, enc_w, enc_b, att_w)
    if nbr.dtype != torch.int32 or nbr.shape[1] not in (16, 32):
        raise ValueError("nbr must be an int32 [N, 16|32] neighbour table")
    if x.shape[1] * 2 != att_w.shape[0] or att_w.shape[0] != att_w.shape[1]:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)}, att_w {tuple(att_w.shape)}")
    return _LFAFunction.apply(x, _f32c(pos), nbr.contiguous(), enc_w, enc_b, att_w)


# ------------------------------------------------------------------------- index / scatter ops
class _GatherRows(torch.autograd.Function):
    """``tensor[idx_decim]`` of decimate() (pyg_randla_net.py:237) and its backward."""

    @staticmethod
    def forward(ctx, x, idx):
        x = _f32c(x)
        out = torch.empty((idx.numel(), x.shape[1]), dtype=torch.float32, device=x.device)
        _call("b200_gather_rows", _p(x), _p(idx), _p(out), idx.numel(), x.shape[1], _stream())
        ctx.save_for_backward(idx)
        ctx.n_src = x.shape[0]
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        gx = torch.zeros((ctx.n_src, grad_out.shape[1]), dtype=torch.float32, device=grad_out.device)
        _call("b200_scatter_rows_add", _p(grad_out), _p(idx), _p(gx), idx.numel(), grad_out.shape[1], _stream())
        return gx, None


MAX_DRAW_KEPT = 25600  # b200_decimation_draw: kept points per cloud that fit its shared-memory sort


def decimation_draw(ptr: Tensor, new_ptr: Tensor, max_kept: int, total_kept: int, seed: int, counter: Optional[Tensor],
                    salt: int) -> Tensor:
    """One launch for ``decimation_indices`` (pyg_randla_net.py:192-231): per cloud a uniformly random ordered
    subset of ``new_ptr[b+1] - new_ptr[b]`` points, as int64 indices into the batch.  ``counter`` (device int64)
    selects the draw; advance it with :func:`counter_add`."""
    _need_cuda(ptr, new_ptr)
    idx = torch.empty(total_kept, dtype=torch.int64, device=ptr.device)
    _call("b200_decimation_draw", _p(ptr), _p(new_ptr), ptr.numel() - 1, max_kept, seed & 0xFFFFFFFFFFFFFFFF, _p(counter),
          salt, _p(idx), _stream())
    return idx


def counter_add(counter: Tensor, delta: int = 1) -> None:
    _call("b200_counter_add", _p(counter), delta, _stream())


def gather_rows(x: Tensor, idx: Tensor) -> Tensor:
    _need_cuda(x, idx)
    if idx.dtype != torch.int64:
        idx = idx.to(torch.int64)
    return _GatherRows.apply(x, idx.contiguous())


class _KnnInterpolate(torch.autograd.Function):
    """Weighting tail of ``knn_interpolate`` (pyg_randla_net.py:250; models/model.py:90-98)."""

    @staticmethod
    def forward(ctx, x, nbr, dist2, k):
        x = _f32c(x)
        ny, c = nbr.shape[0], x.shape[1]
        out = torch.empty((ny, c), dtype=torch.float32, device=x.device)
        _call("b200_knn_interp_fwd", _p(x), _p(nbr), _p(dist2), _p(out), ny, c, k, nbr.shape[1], c, _stream())
        ctx.save_for_backward(nbr, dist2)
        ctx.k, ctx.nx = k, x.shape[0]
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        nbr, dist2 = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        c = grad_out.shape[1]
        gx = torch.zeros((ctx.nx, c), dtype=torch.float32, device=grad_out.device)
        _call("b200_knn_interp_bwd", _p(grad_out), c, _p(nbr), _p(dist2), _p(gx), nbr.shape[0], c, ctx.k,
                                             nbr.shape[1], _stream())
        return gx, None, None, None


def knn_interpolate_from_table(x: Tensor, nbr: Tensor, dist2: Tensor, k: int) -> Tensor:
    _need_cuda(x, nbr, dist2)
    return _KnnInterpolate.apply(x, nbr, dist2, k)


# ---------------------------------------------------------------------------- per-point layers
class _Linear(torch.autograd.Function):
    """``[a1 | a2] @ W^T + b`` (torch.nn.Linear / PyG Linear inside SharedMLP, pyg_randla_net.py:42,53,97-109;
    the concatenation is FPModule's ``torch.cat([x, x_skip], dim=1)``, :251).  Optionally returns the
    per-column (sum, sum of squares) in fp64 for the BatchNorm that follows."""

    @staticmethod
    def forward(ctx, a1, a2, w, b, want_stats):
        ctx.set_materialize_grads(False)  # no zero-filled fp64 "gradient" for the statistics output
        w_param, b_param = w, b
        a1, w = _f32c(a1), _f32c(w)
        a2 = _f32c(a2) if a2 is not None else None
        n, c1 = a1.shape
        c2 = a2.shape[1] if a2 is not None else 0
        cout = w.shape[0]
        if w.shape[1] != c1 + c2:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({n}x{c1 + c2} and {w.shape[1]}x{cout})")
        y = torch.empty((n, cout), dtype=torch.float32, device=a1.device)
        stats = None
        if want_stats:  # per-row-tile partial sums [P, 2*cout]; written by the kernel, summed by bn_finalize
            parts = int(_lib.load().b200_linear_fwd_num_stat_partials(n, c1, c2, cout))
            stats = torch.empty((max(parts, 1), 2 * cout), dtype=torch.float64, device=a1.device)
        _call("b200_linear_fwd", _p(a1), c1, c1, _p(a2), c2, c2, _p(w), _p(b), _p(y), n, cout, _p(stats), _stream())
        ctx.save_for_backward(a1, a2, w)
        ctx.has_bias = b is not None
        ctx.w_param, ctx.b_param = w_param, b_param
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y, None

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y, _grad_stats):
        if grad_y is None:
            return None, None, None, None, None
        a1, a2, w = ctx.saved_tensors
        grad_y = _f32c(grad_y)
        n, c1 = a1.shape
        c2 = a2.shape[1] if a2 is not None else 0
        cout = w.shape[0]
        lib = _lib.load()
        ga1 = torch.empty_like(a1) if ctx.needs_input_grad[0] else None
        ga2 = torch.empty_like(a2) if (a2 is not None and ctx.needs_input_grad[1]) else None
        want_input = ga1 is not None or ga2 is not None
        want_weight = ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3])

        def input_grad():
            nb = int(lib.b200_linear_bwd_input_workspace_bytes(n, c1, c2, cout))
            wsi = torch.empty(nb, dtype=torch.uint8, device=w.device) if nb else None
            _call("b200_linear_bwd_input", _p(grad_y), _p(w), _p(ga1), c1, c1, _p(ga2), c2, c2, _p(wsi), nb, n, cout,
                  _stream())

        fork = want_input and want_weight and _fork_backward()
        if want_input and not fork:
            input_grad()
        gw = gb = None
        if want_weight:
            dw = _direct_grad(ctx.w_param)
            db = _direct_grad(ctx.b_param) if ctx.has_bias else None
            direct = dw is not None and (not ctx.has_bias or db is not None)
            if direct:  # the kernels accumulate: write straight into the parameters' .grad
                gw_buf, gb_buf = dw, db
            else:
                gw_buf = gw = torch.zeros_like(w)
                gb_buf = gb = torch.zeros(cout, dtype=torch.float32, device=w.device) if ctx.has_bias else None
            nbytes = int(lib.b200_linear_bwd_weight_workspace_bytes(n, c1, c2, cout, 1 if ctx.has_bias else 0))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=w.device) if nbytes else None
            if fork:
                # every buffer both branches touch was allocated (and zero-filled) on the main stream above and stays
                # referenced until the join below, so the caching allocator cannot hand it out in between
                main, side = torch.cuda.current_stream(), _side_stream(w.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    _call("b200_linear_bwd_weight", _p(grad_y), _p(a1), c1, c1, _p(a2), c2, c2, _p(gw_buf), _p(gb_buf),
                          _p(ws), nbytes, n, cout, _stream())
                input_grad()
                main.wait_stream(side)
            else:
                _call("b200_linear_bwd_weight", _p(grad_y), _p(a1), c1, c1, _p(a2), c2, c2, _p(gw_buf), _p(gb_buf),
                      _p(ws), nbytes, n, cout, _stream())
        return ga1, ga2, gw, gb, None


def linear(a1: Tensor, w: Tensor, b: Optional[Tensor] = None, a2: Optional[Tensor] = None, want_stats: bool = False):
    _need_cuda(a1, a2, w, b)
    y, stats = _Linear.apply(a1, a2, w, b, want_stats)
    return (y, stats) if want_stats else y


class _BNAct(torch.autograd.Function):
    """``act(BN(y1) [+ BN(y2)])``: BatchNorm1d(momentum .01, eps 1e-6) + LeakyReLU(.2) of SharedMLP
    (pyg_randla_net.py:94-109) and the residual tail ``lrelu(mlp2(x) + shortcut(x))`` (:186-187).

    ``stats*`` are the fp64 column sums from :class:`_Linear` (training) or ``None`` (eval: running
    statistics).  Running buffers are updated in place like ``torch.nn.BatchNorm1d``.
    """

    @staticmethod
    def forward(ctx, y1, stats1, g1, b1, rm1, rv1, nbt1, y2, stats2, g2, b2, rm2, rv2, nbt2, slope, momentum, eps):
        lib = _lib.load()
        n, c = y1.shape
        dev = y1.device
        training = stats1 is not None

        def finalize(stats, g, b, rm, rv, nbt):
            buf = torch.empty((4, c), dtype=torch.float32, device=dev)  # scale, shift, mean, invstd
            parts = stats.shape[0] if stats is not None else 0
            _call("b200_bn_finalize", _p(stats), parts, n, _p(g), _p(b), _p(rm), _p(rv), _p(nbt), momentum, eps,
                                      _p(buf[0]), _p(buf[1]), _p(buf[2]), _p(buf[3]), c, _stream())
            return buf

        params = (g1, b1, g2, b2)
        g1, b1 = _f32c(g1), _f32c(b1)
        y1 = _f32c(y1)
        aff1 = finalize(stats1, g1, b1, rm1, rv1, nbt1)
        aff2 = None
        if y2 is not None:
            g2, b2 = _f32c(g2), _f32c(b2)
            y2 = _f32c(y2)
            aff2 = finalize(stats2, g2, b2, rm2, rv2, nbt2)
        out = torch.empty_like(y1)
        _call("b200_affine_act_fwd", _p(y1), _p(aff1[0]), _p(aff1[1]), _p(y2),
                                     _p(aff2[0]) if aff2 is not None else None,
                                     _p(aff2[1]) if aff2 is not None else None, slope, _p(out), n, c, _stream())
        ctx.save_for_backward(y1, g1, aff1, y2, g2 if y2 is not None else None, aff2, out)
        ctx.slope, ctx.training = slope, training
        ctx.params = params
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        y1, g1, aff1, y2, g2, aff2, out = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        lib = _lib.load()
        n, c = y1.shape
        dev = y1.device
        nred = 4 * c if y2 is not None else 2 * c
        red = _zeros_scratch(nred + 1, torch.float64, dev)  # consumed inside this call; +1: the grid-barrier counter
        red1 = red[:2 * c]
        red2 = red[2 * c:4 * c] if y2 is not None else None
        gy1 = torch.empty_like(y1)
        gy2 = torch.empty_like(y2) if y2 is not None else None
        tr = ctx.training
        gg1 = gb1 = gg2 = gb2 = None
        pg1 = pb1 = pg2 = pb2 = None  # where the kernel accumulates the BatchNorm affine gradients (train mode)
        if tr:
            pg1, pb1 = _direct_grad(ctx.params[0]), _direct_grad(ctx.params[1])
            if pg1 is None or pb1 is None:  # no usable .grad buffer: zero-filled temporaries handed back to autograd
                tmp = torch.zeros(2 * c, dtype=torch.float32, device=dev)  # (autograd may keep them as .grad: no arena)
                pg1, pb1 = gg1, gb1 = tmp[:c], tmp[c:]
            if y2 is not None:
                pg2, pb2 = _direct_grad(ctx.params[2]), _direct_grad(ctx.params[3])
                if pg2 is None or pb2 is None:
                    tmp = torch.zeros(2 * c, dtype=torch.float32, device=dev)
                    pg2, pb2 = gg2, gb2 = tmp[:c], tmp[c:]
            # reduce + apply in one call (one launch with a grid barrier for the small levels, else the two kernels)
            _call("b200_affine_act_bwd", _p(grad_out), _p(out), ctx.slope,
                  _p(y1), _p(g1), _p(aff1[2]), _p(aff1[3]), _p(red1), _p(gy1), _p(pg1), _p(pb1),
                  _p(y2), _p(g2), _p(aff2[2]) if y2 is not None else None, _p(aff2[3]) if y2 is not None else None,
                  _p(red2), _p(gy2), _p(pg2), _p(pb2), _p(red[nred:]), n, c, _stream())
            return (gy1, None, gg1, gb1, None, None, None, gy2, None, gg2, gb2, None, None, None, None, None, None)
        _call("b200_affine_act_bwd_reduce", _p(grad_out), _p(out), ctx.slope, _p(y1), _p(aff1[2]), _p(aff1[3]), _p(red1),
                                            _p(y2), _p(aff2[2]) if y2 is not None else None,
                                            _p(aff2[3]) if y2 is not None else None, _p(red2), n, c, _stream())
        _call("b200_affine_act_bwd_apply",
            _p(grad_out), _p(out), ctx.slope,
            _p(y1), _p(g1), _p(aff1[2]), _p(aff1[3]), None, _p(aff1[0]), _p(gy1), None, None,
            _p(y2), _p(g2), _p(aff2[2]) if y2 is not None else None, _p(aff2[3]) if y2 is not None else None,
            None, _p(aff2[0]) if y2 is not None else None, _p(gy2), None, None,
            n, c, _stream())
        if not tr:  # eval-mode statistics: the same sums are the affine gradients
            gg1, gb1 = red1[c:].float(), red1[:c].float()
            if y2 is not None:
                gg2, gb2 = red2[c:].float(), red2[:c].float()
        return (gy1, None, gg1, gb1, None, None, None, gy2, None, gg2, gb2, None, None, None, None, None, None)


def bn_act(y: Tensor, stats: Optional[Tensor], bn: torch.nn.BatchNorm1d, slope: float,
           y2: Optional[Tensor] = None, stats2: Optional[Tensor] = None, bn2: Optional[torch.nn.BatchNorm1d] = None) -> Tensor:
    """BatchNorm (+ second normalised branch) + LeakyReLU(slope) (slope=1.0: no activation)."""
    _need_cuda(y, y2)
    training = stats is not None
    if training and y.shape[0] <= 1:
        # same failure as torch.nn.BatchNorm1d in training mode
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(y.shape)}")

    def buffers(m):
        if m.running_mean is None:
            raise RuntimeError("BatchNorm without running statistics is not supported")
        return m.running_mean, m.running_var

    rm1, rv1 = buffers(bn)
    args2 = (None, None, None, None, None, None, None)
    if y2 is not None:
        rm2, rv2 = buffers(bn2)
        args2 = (y2, stats2, bn2.weight, bn2.bias, rm2, rv2, bn2.num_batches_tracked)
    return _BNAct.apply(y, stats, bn.weight, bn.bias, rm1, rv1, bn.num_batches_tracked, *args2, float(slope),
                        float(bn.momentum), float(bn.eps))


# ------------------------------------------------------------------------------ loss
class _CrossEntropy(torch.autograd.Function):
    """``torch.nn.CrossEntropyLoss(weight, ignore_index, reduction="mean")`` (models/model.py:117-118)."""

    @staticmethod
    def forward(ctx, logits, target, weight, ignore_index):
        logits = _f32c(logits)
        target = target.contiguous()
        weight = _f32c(weight) if weight is not None else None
        n, c = logits.shape
        scratch = _zeros_scratch(4, torch.float64, logits.device)  # [sum, weight sum] fp64 + a uint32 CTA counter
        out = torch.empty(2, dtype=torch.float32, device=logits.device)
        _call("b200_cross_entropy_fwd", _p(logits), _p(target), _p(weight), n, c, int(ignore_index), _p(scratch[:2]),
              _p(scratch[2:]), _p(out), _stream())
        ctx.save_for_backward(logits, target, weight, out)
        ctx.ignore_index = int(ignore_index)
        return out[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loss):
        logits, target, weight, out = ctx.saved_tensors
        n, c = logits.shape
        grad = torch.empty_like(logits)
        g = _f32c(grad_loss).reshape(1)
        _call("b200_cross_entropy_bwd", _p(logits), _p(target), _p(weight), n, c, ctx.ignore_index, _p(out), _p(g),
              _p(grad), _stream())
        return grad, None, None, None


CROSS_ENTROPY_MAX_CLASSES = 32


def cross_entropy(logits: Tensor, target: Tensor, weight: Optional[Tensor] = None, ignore_index: int = -100) -> Tensor:
    """Mean cross-entropy over the rows whose target is not ``ignore_index`` (class-weighted mean if ``weight``)."""
    _need_cuda(logits, target, weight)
    if logits.dim() != 2 or target.shape != logits.shape[:1] or target.dtype != torch.int64:
        raise ValueError(f"cross_entropy expects [N, C] logits and int64 [N] targets, got {tuple(logits.shape)}, "
                         f"{tuple(target.shape)} {target.dtype}")
    if weight is not None and weight.shape != (logits.shape[1],):
        raise ValueError("weight must have one entry per class")
    return _CrossEntropy.apply(logits, target, weight, ignore_index)


# ------------------------------------------------------------------------------ sliding-window stitch
def stitch_scatter_sum(logits: Tensor, idx: Tensor, nb_points: int) -> Tensor:
    """``scatter_sum(logits, idx, out=zeros(nb_points, C))`` (``interpolation.py:113-116``), summing the predictions of
    one point in input order (bit-identical to the reference's CPU scatter; deterministic)."""
    _need_cuda(logits, idx)
    logits = _f32c(logits)
    idx = idx.to(torch.int64).contiguous()
    m, c = logits.shape
    if idx.numel() != m:
        raise ValueError(f"stitch_scatter_sum: {idx.numel()} indices for {m} rows")
    if m and (int(idx.min()) < 0 or int(idx.max()) >= nb_points):
        raise IndexError(f"stitch_scatter_sum: index out of range for {nb_points} points")
    out = torch.zeros(nb_points, c, dtype=torch.float32, device=logits.device)
    if m == 0:
        return out
    sorted_idx, order = torch.sort(idx, stable=True)
    _call("b200_stitch_segment_sum", _p(logits), _p(order), _p(sorted_idx), _p(out), m, c, nb_points, _stream())
    return out


def stitch_finalize(reduced: Tensor, idx: Tensor, want_logits: bool = True):
    """``reduced[idx]`` + softmax + argmax + Shannon entropy (``interpolation.py:121,142-166``) in one kernel.
    Returns ``(logits | None, probas, preds int64, entropy)``, each with ``len(idx)`` rows."""
    _need_cuda(reduced, idx)
    reduced = _f32c(reduced)
    idx = idx.to(torch.int64).contiguous()
    m, c = idx.numel(), reduced.shape[1]
    dev = reduced.device
    logits = torch.empty(m, c, dtype=torch.float32, device=dev) if want_logits else None
    probas = torch.empty(m, c, dtype=torch.float32, device=dev)
    preds = torch.empty(m, dtype=torch.int64, device=dev)
    entropy = torch.empty(m, dtype=torch.float32, device=dev)
    if m:
        _call("b200_stitch_finalize", _p(reduced), _p(idx), _p(logits), _p(probas), _p(preds), _p(entropy), m, c, _stream())
    return logits, probas, preds, entropy
