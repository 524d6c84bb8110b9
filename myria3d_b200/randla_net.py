"""B200-native RandLA-Net: drop-in for ``PyGRandLANet``
(``myria3d/models/modules/pyg_randla_net.py:22-253``).

Same constructor, same ``forward(x, pos, batch, ptr)``, same parameter / buffer names (so
``load_state_dict(strict=True)`` accepts myria3d checkpoints) -- but the forward
and backward run on hand-written sm_100a kernels through ``myria3d_b200.ops`` instead of
torch_geometric / torch_cluster / torch_scatter.  CUDA only: there is no CPU fallback.
"""
from __future__ import annotations

from numbers import Number
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops
from .ops import BN_EPS, BN_MOMENTUM, LRELU_SLOPE


class PyGBatchNorm(nn.Module):
    """Name-compatible stand-in for ``torch_geometric.nn.norm.BatchNorm`` (keys ``norms.N.module.*``)."""

    def __init__(self, channels: int):
        super().__init__()
        self.module = nn.BatchNorm1d(channels, eps=BN_EPS, momentum=BN_MOMENTUM)


class SharedMLP(nn.Module):
    """``SharedMLP`` of pyg_randla_net.py:97-109 (PyG ``MLP`` with ``plain_last=False``): per layer
    Linear -> BatchNorm1d(momentum .01, eps 1e-6) -> LeakyReLU(.2) -> dropout.  Parameters live in
    ``lins`` / ``norms`` exactly as in PyG; the computation is ``ops.linear`` + ``ops.bn_act``."""

    def __init__(self, channels: Sequence[int], dropout=0.0, act: bool = True, norm: bool = True, bias: bool = True):
        super().__init__()
        n = len(channels) - 1
        self.dropout = list(dropout) if isinstance(dropout, (list, tuple)) else [float(dropout)] * n
        self.act = act
        self.lins = nn.ModuleList([nn.Linear(channels[i], channels[i + 1], bias=bias) for i in range(n)])
        self.norms = nn.ModuleList([PyGBatchNorm(channels[i + 1]) if norm else nn.Identity() for i in range(n)])
        self.injected_masks: Optional[List[Optional[Tensor]]] = None  # parity harness only

    def forward(self, x: Tensor, x2: Optional[Tensor] = None) -> Tensor:
        """``x2`` (optional) is concatenated to ``x`` along channels without materialising the cat."""
        for i, (lin, norm) in enumerate(zip(self.lins, self.norms)):
            second = x2 if i == 0 else None
            if isinstance(norm, PyGBatchNorm):
                if self.training:
                    y, stats = ops.linear(x, lin.weight, lin.bias, a2=second, want_stats=True)
                else:
                    y, stats = ops.linear(x, lin.weight, lin.bias, a2=second), None
                x = ops.bn_act(y, stats, norm.module, LRELU_SLOPE if self.act else 1.0)
            else:
                x = ops.linear(x, lin.weight, lin.bias, a2=second)
                if self.act:
                    x = F.leaky_relu(x, LRELU_SLOPE)
            p = self.dropout[i]
            if self.training and self.injected_masks is not None and self.injected_masks[i] is not None:
                x = x * self.injected_masks[i]
            elif p > 0.0:
                x = F.dropout(x, p=p, training=self.training)
        return x


def fold_encoder(enc: SharedMLP, moments: Optional[Tensor], num_edges: int, training: bool):
    """Fold ``mlp_encoder`` (Linear(10->h) + BatchNorm, pyg_randla_net.py:117,144) into an affine map
    of ``q = (p_i, p_j, |p_j - p_i|)``.  Plain-torch SPECIFICATION of ``ops.encoder_fold`` (the kernel
    ``b200_encoder_fold_fwd/bwd`` is what the network runs); kept for the tests that pin the algebra.

    The reference feeds ``r = [p_i, p_j, p_j - p_i, dist]`` (:143): ``W r = W~ q`` with
    ``W~ = [W_a - W_c, W_b + W_c, w_d]``.  In training the BatchNorm statistics over all E edges are
    ``mean = W~ mu_q + b`` and ``var = diag(W~ C_q W~^T)`` from the fp64 edge moments (SURVEY.md
    App. D-7); everything here is differentiable torch on [h, 7] tensors, so autograd returns the
    exact train-mode BatchNorm gradients for W, b, gamma, beta.  Running statistics are updated as
    ``torch.nn.BatchNorm1d`` does (unbiased variance, momentum .01).
    """
    lin, bn = enc.lins[0], enc.norms[0].module
    w, b = lin.weight, lin.bias
    wq = torch.cat([w[:, 0:3] - w[:, 6:9], w[:, 3:6] + w[:, 6:9], w[:, 9:10]], dim=1).double()  # [h, 7]
    bd = b.double() if b is not None else torch.zeros(w.shape[0], dtype=torch.float64, device=w.device)
    if training:
        if num_edges <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size [{num_edges}, 10]")
        e = moments[0]
        mu = moments[1:8] / e
        cov = moments[8:57].view(7, 7) / e - torch.outer(mu, mu)
        mean = wq @ mu + bd
        var = ((wq @ cov) * wq).sum(dim=1).clamp_min(0.0)
        with torch.no_grad():
            m = bn.momentum
            bn.running_mean.mul_(1.0 - m).add_(mean.float(), alpha=m)
            bn.running_var.mul_(1.0 - m).add_((var * (num_edges / (num_edges - 1.0))).float(), alpha=m)
            bn.num_batches_tracked += 1
    else:
        mean = bn.running_mean.double()
        var = bn.running_var.double()
    scale = bn.weight.double() * torch.rsqrt(var + bn.eps)
    enc_w = (wq * scale[:, None]).float()
    enc_b = ((bd - mean) * scale + bn.bias.double()).float()
    return enc_w, enc_b


class _Level:
    """Host + device bookkeeping of one resolution level (clouds stay contiguous)."""

    def __init__(self, ptr_host: List[int], device: torch.device, ptr_dev: Optional[Tensor] = None):
        self.ptr_host = ptr_host
        self.n = ptr_host[-1]
        sizes = [ptr_host[i + 1] - ptr_host[i] for i in range(len(ptr_host) - 1)]
        self.sizes = sizes
        self.max_n = max(sizes) if sizes else 0
        self.ptr = ptr_dev if ptr_dev is not None else torch.tensor(ptr_host, dtype=torch.int64, device=device)

        self._decim_cache = None

    def num_edges(self, k: int) -> int:
        return sum(n * min(k, n) for n in self.sizes)

    def decimation_tables(self, new_ptr_host: Sequence[int]):
        """(cloud id in the high key bits per point, positions kept per cloud, random bits per key) for
        :func:`fused_decimation_indices`."""
        if self._decim_cache is None:
            dev = self.ptr.device
            sizes = torch.tensor(self.sizes, dtype=torch.int64)
            cloud = torch.repeat_interleave(torch.arange(len(self.sizes), dtype=torch.int64), sizes)
            take = torch.cat([torch.arange(self.ptr_host[b], self.ptr_host[b] + (new_ptr_host[b + 1] - new_ptr_host[b]),
                                           dtype=torch.int64) for b in range(len(self.sizes))])
            bits = max(1, (len(self.sizes) - 1).bit_length())
            if bits <= 5:  # <= 32 clouds: (cloud id, >= 26 random bits) fits a non-negative int32 -> half the radix passes
                self._decim_cache = ((cloud << (31 - bits)).to(torch.int32).to(dev), take.to(dev), 31 - bits)
            else:
                self._decim_cache = ((cloud << 32).to(dev), take.to(dev), 31)
        return self._decim_cache


class LocalFeatureAggregation(nn.Module):
    """pyg_randla_net.py:112-152: LocSE + attentive pooling (one fused kernel) + post-attention SharedMLP."""

    def __init__(self, channels: int):
        super().__init__()
        self.mlp_encoder = SharedMLP([10, channels // 2])
        self.mlp_attention = SharedMLP([channels, channels], bias=False, act=False, norm=False)
        self.mlp_post_attention = SharedMLP([channels, channels])

    def forward(self, x: Tensor, pos: Tensor, nbr: Tensor, moments: Optional[Tensor], num_edges: int) -> Tensor:
        enc = self.mlp_encoder
        enc_w, enc_b = ops.encoder_fold(enc.lins[0], enc.norms[0].module, moments, num_edges, self.training)
        pooled = ops.lfa_attentive_pool(x, pos, nbr, enc_w, enc_b, self.mlp_attention.lins[0].weight)
        return self.mlp_post_attention(pooled)


class DilatedResidualBlock(nn.Module):
    """pyg_randla_net.py:155-189."""

    def __init__(self, num_neighbors: int, d_in: int, d_out: int):
        super().__init__()
        self.num_neighbors = num_neighbors
        self.d_in = d_in
        self.d_out = d_out
        self.mlp1 = SharedMLP([d_in, d_out // 8])
        self.shortcut = SharedMLP([d_in, d_out], act=False)
        self.mlp2 = SharedMLP([d_out // 2, d_out], act=False)
        self.lfa1 = LocalFeatureAggregation(d_out // 4)
        self.lfa2 = LocalFeatureAggregation(d_out // 2)
        self.last_nbr: Optional[Tensor] = None

    def graph(self, pos: Tensor, level: _Level):
        """knn_graph(pos, k, batch, loop=True) (:180) once per block, shared by both LFAs, + the fp64 edge moments the
        folded encoder BatchNorm needs in training."""
        k = self.num_neighbors
        nbr, _ = ops.knn(pos, level.ptr, pos, level.ptr, k, level.max_n, kt=ops.table_width(k), want_dist=False)
        moments = ops.edge_moments(pos, nbr) if self.training else None
        return nbr, moments

    def forward(self, x: Tensor, pos: Tensor, level: _Level, geometry: Optional["_Geometry"] = None, lvl_idx: int = 0) -> Tensor:
        k = self.num_neighbors
        num_edges = level.num_edges(k)
        if geometry is None:
            nbr, moments = self.graph(pos, level)

        sc_lin, sc_bn = self.shortcut.lins[0], self.shortcut.norms[0].module
        m2_lin, m2_bn = self.mlp2.lins[0], self.mlp2.norms[0].module
        if self.training:
            y_sc, st_sc = ops.linear(x, sc_lin.weight, sc_lin.bias, want_stats=True)  # :182
        else:
            y_sc, st_sc = ops.linear(x, sc_lin.weight, sc_lin.bias), None
        h = self.mlp1(x)  # :183
        if geometry is not None:  # built ahead (possibly on a second stream): the shortcut and mlp1 did not need it
            nbr, moments = geometry.graph_of(lvl_idx)
        self.last_nbr = nbr
        h = self.lfa1(h, pos, nbr, moments, num_edges)  # :184
        h = self.lfa2(h, pos, nbr, moments, num_edges)  # :185
        if self.training:
            y2, st2 = ops.linear(h, m2_lin.weight, m2_lin.bias, want_stats=True)  # :186
        else:
            y2, st2 = ops.linear(h, m2_lin.weight, m2_lin.bias), None
        # lrelu(BN(mlp2) + BN(shortcut)) in one pass (:187)
        return ops.bn_act(y2, st2, m2_bn, LRELU_SLOPE, y2=y_sc, stats2=st_sc, bn2=sc_bn)


def decimation_sizes(ptr_host: Sequence[int], decimation_factor: Number) -> List[int]:
    """New ``ptr`` after ``decimate`` (pyg_randla_net.py:214-229): ``max(1, floor(n / factor))`` per cloud."""
    if decimation_factor < 1:
        raise ValueError(
            "Argument `decimation_factor` should be higher than (or equal to) "
            f"1 for downsampling. (Current value: {decimation_factor})"
        )
    new_ptr = [int(ptr_host[0])]
    for b in range(len(ptr_host) - 1):
        n = int(ptr_host[b + 1]) - int(ptr_host[b])
        new_ptr.append(new_ptr[-1] + max(1, int(n // decimation_factor)))
    return new_ptr


def decimation_indices(ptr_host: Sequence[int], decimation_factor: Number, device: torch.device):
    """``decimation_indices`` of pyg_randla_net.py:192-231: the same ``torch.randperm(n_i, device=...)``
    calls in the same (cloud) order as the reference, hence the same permutations under a fixed seed
    on the same device -- but cloud sizes come from the host copy of ``ptr`` (no device sync)."""
    new_ptr = decimation_sizes(ptr_host, decimation_factor)
    parts = []
    for b in range(len(ptr_host) - 1):
        n = int(ptr_host[b + 1]) - int(ptr_host[b])
        nd = new_ptr[b + 1] - new_ptr[b]
        parts.append(int(ptr_host[b]) + torch.randperm(n, device=device)[:nd])
    return torch.cat(parts, dim=0), new_ptr


def fused_decimation_indices(level: "_Level", new_ptr_host: Sequence[int]) -> Tensor:
    """Same distribution as :func:`decimation_indices` -- per cloud, a uniformly random ORDERED subset of
    ``max(1, n // decimation)`` points -- drawn for ALL clouds at once: one random key per point, one sort of
    (cloud id, key), one gather of each cloud's leading positions.  ~4 launches per level instead of
    ~7 per cloud (the reference's Python loop); the random stream differs from the
    reference's per-cloud ``randperm`` calls, the statistics do not."""
    shift, take, random_bits = level.decimation_tables(new_ptr_host)
    # ties between two keys of one cloud (expected < 1 pair per 12 800-point cloud at 27 bits) fall back to index order
    keys = torch.randint(0, 2 ** random_bits - 1, (level.n,), dtype=shift.dtype, device=shift.device)
    order = torch.argsort(keys + shift)
    return order[take]


class FPModule(nn.Module):
    """pyg_randla_net.py:241-253: k-NN (k=1) inverse-distance upsampling + skip concat + SharedMLP."""

    def __init__(self, k: int, net: nn.Module):
        super().__init__()
        self.k = k
        self.nn = net

    def table(self, pos: Tensor, level: _Level, pos_skip: Tensor, level_skip: _Level):
        """The k nearest coarse points of every fine point and their squared distances (knn_interpolate's search)."""
        return ops.knn(pos, level.ptr, pos_skip, level_skip.ptr, self.k, level_skip.max_n, kt=self.k,
                       max_points_per_cloud=level.max_n)

    def forward(self, x: Tensor, pos: Tensor, level: _Level, x_skip: Tensor, pos_skip: Tensor, level_skip: _Level,
                table=None) -> Tensor:
        nbr, dist2 = table if table is not None else self.table(pos, level, pos_skip, level_skip)
        xi = ops.knn_interpolate_from_table(x, nbr, dist2, self.k)  # :250
        return self.nn(xi, x_skip)  # cat + SharedMLP (:251-252) without materialising the cat


class _Geometry:
    """Everything of a forward pass that depends on the POSITIONS only: per level the kNN graph, its edge moments, the
    decimation draw, the decimated positions and the decoder's interpolation tables.  None of it needs the features, so
    :meth:`B200RandLANet._geometry` computes it ahead of the encoder -- while a CUDA graph is being captured on a second
    stream, as a parallel branch that runs next to the feature kernels (kNN is latency-bound, the Linear / BatchNorm
    passes are HBM-bound).  ``ready[l]`` is recorded on that stream once level ``l`` is complete."""

    def __init__(self):
        self.pos: List[Tensor] = []
        self.graphs: list = []   # level -> (nbr, moments)
        self.idx: List[Tensor] = []    # level -> decimation indices into it
        self.tables: list = []   # level l -> (nbr, dist2) of the l+1 -> l interpolation
        self.ready: list = []    # level -> torch.cuda.Event (graph of the level complete) or None
        self.moved: list = []    # level -> torch.cuda.Event (its draw and the next level's positions complete) or None
        self.done = None

    def _wait(self, ev) -> None:
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def graph_of(self, lvl_idx: int):
        self._wait(self.ready[lvl_idx])
        return self.graphs[lvl_idx]

    def after_level(self, lvl_idx: int):
        """(decimation indices of level ``lvl_idx``, positions of level ``lvl_idx + 1``)."""
        self._wait(self.moved[lvl_idx])
        return self.idx[lvl_idx], self.pos[lvl_idx + 1]

    def table_of(self, lvl_idx: int):
        self._wait(self.done)
        return self.tables[lvl_idx]


class B200RandLANet(nn.Module):
    """Drop-in for ``PyGRandLANet`` (pyg_randla_net.py:22-88) on B200.

    Registered in ``myria3d_b200.model.MODEL_ZOO``; select it with
    ``model.neural_net_class_name=B200RandLANet`` in myria3d's Hydra config.
    """

    def __init__(self, num_features: int, num_classes: int, decimation: int = 4, num_neighbors: int = 16,
                 return_logits: bool = False):
        super().__init__()
        ops.table_width(num_neighbors)  # validates num_neighbors <= 32
        self.decimation = decimation
        self.return_logits = return_logits
        d_bottleneck = max(32, num_classes, num_features)  # :40

        self.fc0 = nn.Linear(num_features, d_bottleneck)
        self.block1 = DilatedResidualBlock(num_neighbors, d_bottleneck, 32)
        self.block2 = DilatedResidualBlock(num_neighbors, 32, 128)
        self.block3 = DilatedResidualBlock(num_neighbors, 128, 256)
        self.block4 = DilatedResidualBlock(num_neighbors, 256, 512)
        self.mlp_summit = SharedMLP([512, 512])
        self.fp4 = FPModule(1, SharedMLP([512 + 256, 256]))
        self.fp3 = FPModule(1, SharedMLP([256 + 128, 128]))
        self.fp2 = FPModule(1, SharedMLP([128 + 32, 32]))
        self.fp1 = FPModule(1, SharedMLP([32 + 32, d_bottleneck]))
        self.mlp_classif = SharedMLP([d_bottleneck, 64, 32], dropout=[0.0, 0.5])
        self.fc_classif = nn.Linear(32, num_classes)

        # host-side caches / static-shape hooks (see myria3d_b200.graphed.GraphedTrainStep)
        self._level_cache: Dict[tuple, List[_Level]] = {}
        self.static_ptr_host: Optional[List[int]] = None  # skips the ptr device->host read when set
        # "reference": per-cloud torch.randperm calls exactly like pyg_randla_net.py:219-229 (same subsets as
        # the reference under a fixed seed on the same device); "fused": one batched draw per level.
        self.decimation_rng = "reference"
        self._draw_counter: Optional[Tensor] = None  # device int64: index of the next fused draw
        self._draw_seed = 0
        # parity harness hooks (never set in production)
        self.injected_decimation_idx: Optional[List[Tensor]] = None
        self.last_decimation_idx: List[Tensor] = []
        self.keep_stages = False
        self.stages: Dict[str, Tensor] = {}

    def levels_for(self, ptr_host: Sequence[int], device: torch.device) -> List["_Level"]:
        """The 5 resolution levels (ptr on host and device) of a batch layout.  Cloud sizes after each
        decimation are deterministic (max(1, n // decimation)), so they are computed once per layout
        and cached: no per-step host->device copies, and a static layout can be CUDA-graph captured."""
        key = (tuple(int(v) for v in ptr_host), str(device))
        levels = self._level_cache.get(key)
        if levels is None:
            levels = [_Level(list(key[0]), device)]
            for _ in range(4):
                levels.append(_Level(decimation_sizes(levels[-1].ptr_host, self.decimation), device))
            if len(self._level_cache) >= 16:
                self._level_cache.pop(next(iter(self._level_cache)))
            self._level_cache[key] = levels
        return levels

    def draw_decimation(self, levels: List["_Level"], lvl_idx: int) -> Tensor:
        """``"fused"``: all clouds of the level in ONE kernel launch (``b200_decimation_draw``: Philox keys, radix select,
        shared-memory sort; graph-capturable, the stream advances through a device counter); ``"reference"``: the
        reference's per-cloud ``torch.randperm`` calls (same subsets as the reference under a fixed seed)."""
        if self.decimation_rng == "fused":
            lvl, nxt = levels[lvl_idx], levels[lvl_idx + 1]
            if nxt.max_n > ops.MAX_DRAW_KEPT:  # > 102 400 points per cloud: batched torch draw (random keys + sort)
                return fused_decimation_indices(lvl, nxt.ptr_host)
            if self._draw_counter is None or self._draw_counter.device != lvl.ptr.device:
                self._draw_counter = torch.zeros(1, dtype=torch.int64, device=lvl.ptr.device)
                self._draw_seed = int(torch.initial_seed())
            idx = ops.decimation_draw(lvl.ptr, nxt.ptr, nxt.max_n, nxt.n, self._draw_seed, self._draw_counter, lvl_idx)
            if lvl_idx == 3:  # the last draw of a forward pass: the next pass gets fresh permutations
                ops.counter_add(self._draw_counter, 1)
            return idx
        return decimation_indices(levels[lvl_idx].ptr_host, self.decimation, levels[lvl_idx].ptr.device)[0]

    def _decimation_idx(self, levels: List["_Level"], lvl_idx: int, device: torch.device) -> Tensor:
        """decimation_indices() of pyg_randla_net.py:192-231 for one level (or the parity harness's injected ones)."""
        if self.injected_decimation_idx is not None:
            idx = self.injected_decimation_idx[lvl_idx].to(device=device, dtype=torch.int64)
        else:
            idx = self.draw_decimation(levels, lvl_idx)
        self.last_decimation_idx.append(idx)
        return idx

    def _geometry(self, pos: Tensor, levels: List["_Level"]) -> _Geometry:
        """The position-only half of the forward pass (see :class:`_Geometry`): same calls in the same order as the
        interleaved reference code (:58-79), so the random stream of the draws is consumed identically."""
        g = _Geometry()
        blocks = (self.block1, self.block2, self.block3, self.block4)
        fork = ops.fork_enabled()
        if fork:
            main, side = torch.cuda.current_stream(), ops.side_stream(pos.device)
            side.wait_stream(main)
            ctx = torch.cuda.stream(side)
        else:
            import contextlib
            ctx = contextlib.nullcontext()

        def mark():
            if fork:
                ev = torch.cuda.Event()
                ev.record(side)
                return ev
            return None

        with ctx:
            g.pos.append(pos)
            for l, block in enumerate(blocks):
                g.graphs.append(block.graph(g.pos[l], levels[l]))  # knn_graph (:180)
                g.ready.append(mark())
                idx = self._decimation_idx(levels, l, pos.device)   # decimate (:59-68)
                g.idx.append(idx)
                g.pos.append(ops.gather_rows(g.pos[l], idx))
                g.moved.append(mark())
            fps = (self.fp1, self.fp2, self.fp3, self.fp4)
            for l, fp in enumerate(fps):  # level l+1 -> level l (knn_interpolate's search, :248-250)
                g.tables.append(fp.table(g.pos[l + 1], levels[l + 1], g.pos[l], levels[l]))
            g.done = mark()
        return g

    def forward(self, x: Optional[Tensor], pos: Tensor, batch: Optional[Tensor], ptr: Tensor) -> Tensor:
        if not pos.is_cuda:
            raise RuntimeError("B200RandLANet runs on a CUDA (B200) device only; there is no CPU fallback")
        x = x if x is not None else pos  # :56
        pos = pos.float().contiguous()
        ops.reset_scratch(pos.device)  # one memset for all the small zero-initialised buffers of the last pass
        if self.decimation < 1:
            decimation_sizes([0, 1], self.decimation)  # raises the reference's ValueError
        if self.static_ptr_host is not None:
            ptr_host = self.static_ptr_host
        else:
            ptr_host = [int(v) for v in ptr.tolist()]  # the only device->host sync of the forward
        levels = self.levels_for(ptr_host, pos.device)
        lvl0, lvl1, lvl2, lvl3, lvl4 = levels
        self.last_decimation_idx = []
        self.stages = {}

        def keep(name, t):
            if self.keep_stages:
                self.stages[name] = t

        geo = self._geometry(pos, levels)
        h0 = ops.linear(x, self.fc0.weight, self.fc0.bias)  # fc0 (:58)
        b1 = self.block1(h0, pos, lvl0, geo, 0)
        keep("b1", b1)
        idx0, pos1 = geo.after_level(0)
        b1d = ops.gather_rows(b1, idx0)  # decimate (:59)
        b2 = self.block2(b1d, pos1, lvl1, geo, 1)
        keep("b2", b2)
        idx1, pos2 = geo.after_level(1)
        b2d = ops.gather_rows(b2, idx1)  # :62
        b3 = self.block3(b2d, pos2, lvl2, geo, 2)
        keep("b3", b3)
        idx2, pos3 = geo.after_level(2)
        b3d = ops.gather_rows(b3, idx2)  # :65
        b4 = self.block4(b3d, pos3, lvl3, geo, 3)
        keep("b4", b4)
        idx3, pos4 = geo.after_level(3)
        b4d = ops.gather_rows(b4, idx3)  # :68

        summit = self.mlp_summit(b4d)  # :70
        keep("summit", summit)
        fp4 = self.fp4(summit, pos4, lvl4, b3d, pos3, lvl3, geo.table_of(3))  # :76
        fp3 = self.fp3(fp4, pos3, lvl3, b2d, pos2, lvl2, geo.table_of(2))
        fp2 = self.fp2(fp3, pos2, lvl2, b1d, pos1, lvl1, geo.table_of(1))
        fp1 = self.fp1(fp2, pos1, lvl1, b1, pos, lvl0, geo.table_of(0))  # :79
        keep("fp1", fp1)

        h = self.mlp_classif(fp1)  # :81
        logits = ops.linear(h, self.fc_classif.weight, self.fc_classif.bias)  # :82
        if self.return_logits:
            return logits
        return logits.log_softmax(dim=-1)  # :87


class B200Block1Net(nn.Module):
    """``fc0 + block1 + head`` -- the "1 encoder layer" network of BASELINE.json configs[0] (SURVEY.md section 8d,
    config A), the reference's CPU-runnable case: no decimation, no decoder.  Same sub-module names as
    :class:`B200RandLANet` (``fc0``, ``block1``, ``mlp_classif``, ``fc_classif``: a full state dict loads with
    ``strict=False``), same kernels, same ``forward(x, pos, batch, ptr)`` signature."""

    def __init__(self, num_features: int, num_classes: int, num_neighbors: int = 16, return_logits: bool = True):
        super().__init__()
        ops.table_width(num_neighbors)
        d_bottleneck = max(32, num_classes, num_features)  # pyg_randla_net.py:40
        self.fc0 = nn.Linear(num_features, d_bottleneck)
        self.block1 = DilatedResidualBlock(num_neighbors, d_bottleneck, 32)
        self.mlp_classif = SharedMLP([32, 64, 32], dropout=[0.0, 0.5])
        self.fc_classif = nn.Linear(32, num_classes)
        self.return_logits = return_logits

    def forward(self, x: Optional[Tensor], pos: Tensor, batch: Optional[Tensor], ptr: Tensor) -> Tensor:
        if not pos.is_cuda:
            raise RuntimeError("B200Block1Net runs on a CUDA (B200) device only; there is no CPU fallback")
        x = x if x is not None else pos
        pos = pos.float().contiguous()
        ops.reset_scratch(pos.device)
        lvl0 = _Level([int(v) for v in ptr.tolist()], pos.device)
        h0 = ops.linear(x, self.fc0.weight, self.fc0.bias)
        b1 = self.block1(h0, pos, lvl0)
        logits = ops.linear(self.mlp_classif(b1), self.fc_classif.weight, self.fc_classif.bias)
        return logits if self.return_logits else logits.log_softmax(dim=-1)
