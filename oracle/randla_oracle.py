"""CPU restatement of myria3d's RandLA-Net forward (autograd gives the backward).

TEST INFRASTRUCTURE -- never imported by the product package ``myria3d_b200``.
PARITY UNPINNED (see ``oracle/__init__.py``).

Every function cites the reference line it restates; paths are relative to
``/root/reference``.  The third-party semantics (PyG 2.4 ``MLP`` /
``MessagePassing`` / ``softmax`` / ``knn_interpolate``, torch_cluster ``knn``,
torch_scatter ``scatter``) are restated from the published behaviour of the
versions pinned in ``environment.yml:14-22``; only plain ``torch`` CPU ops and
scipy's ``cKDTree`` (the closest available stand-in for torch_cluster's
nanoflann kd-tree) are used, in the same op order as the reference so the fp32
rounding sequence is the reference's.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

LRELU_SLOPE = 0.2  # myria3d/models/modules/pyg_randla_net.py:92
BN_MOMENTUM = 0.01  # pyg_randla_net.py:94
BN_EPS = 1e-6  # pyg_randla_net.py:94


# --------------------------------------------------------------------------- kNN
def _canonical_d2(q: Tensor, p: Tensor) -> Tensor:
    """fp32 squared distance, canonical rounding order ((dx*dx + dy*dy) + dz*dz),
    separate mul / add roundings (no FMA).  q [..., 3], p [..., 3] broadcastable."""
    d = p - q
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return (dx * dx + dy * dy) + dz * dz


def knn_bruteforce(pos_x: Tensor, ptr_x: Sequence[int], pos_y: Tensor, ptr_y: Sequence[int], k: int,
                   chunk: int = 2048) -> Tuple[Tensor, Tensor]:
    """Exact per-cloud k-NN of every y among the x of the same cloud.

    Restates torch_cluster ``knn(x, y, k, batch_x, batch_y)`` (called through
    ``knn_graph`` at pyg_randla_net.py:180 and ``knn_interpolate`` at :250):
    ascending squared Euclidean distance, only ``min(k, n_cloud)`` neighbours
    for small clouds.  Ties are broken towards the LOWER x index (the strict
    ``>`` insertion of torch_cluster's CUDA kernel); the set is what parity is
    asserted on.

    Returns ``nbr`` int64 [Ny, k] (GLOBAL x indices, -1 padded) and ``deg`` int64 [Ny].
    """
    ny = pos_y.shape[0]
    nbr = torch.full((ny, k), -1, dtype=torch.int64)
    deg = torch.zeros(ny, dtype=torch.int64)
    for b in range(len(ptr_x) - 1):
        xs, xe = int(ptr_x[b]), int(ptr_x[b + 1])
        ys, ye = int(ptr_y[b]), int(ptr_y[b + 1])
        n = xe - xs
        kk = min(k, n)
        if n == 0 or ye == ys:
            continue
        px = pos_x[xs:xe]
        for s in range(ys, ye, chunk):
            e = min(ye, s + chunk)
            d2 = _canonical_d2(pos_y[s:e, None, :], px[None, :, :])  # [q, n]
            # stable sort => ties keep the lower index first
            order = torch.sort(d2, dim=1, stable=True).indices[:, :kk]
            nbr[s:e, :kk] = order + xs
            deg[s:e] = kk
    return nbr, deg


def knn_kdtree(pos_x: Tensor, ptr_x: Sequence[int], pos_y: Tensor, ptr_y: Sequence[int], k: int,
               extra: int = 8, workers: int = 1) -> Tuple[Tensor, Tensor]:
    """Same contract as :func:`knn_bruteforce`, via scipy ``cKDTree`` (kd-tree like
    torch_cluster's CPU path, ``num_workers=1`` as at pyg_randla_net.py:180) and an
    fp32 canonical re-rank of ``k+extra`` candidates; falls back to brute force for the
    (vanishingly rare) queries whose candidate list cannot prove the top-k."""
    from scipy.spatial import cKDTree

    ny = pos_y.shape[0]
    nbr = torch.full((ny, k), -1, dtype=torch.int64)
    deg = torch.zeros(ny, dtype=torch.int64)
    for b in range(len(ptr_x) - 1):
        xs, xe = int(ptr_x[b]), int(ptr_x[b + 1])
        ys, ye = int(ptr_y[b]), int(ptr_y[b + 1])
        n = xe - xs
        if n == 0 or ye == ys:
            continue
        kk = min(k, n)
        kc = min(n, kk + extra)
        px = pos_x[xs:xe]
        py = pos_y[ys:ye]
        tree = cKDTree(px.numpy().astype(np.float64))
        _, cand = tree.query(py.numpy().astype(np.float64), k=kc, workers=workers)
        cand = torch.from_numpy(np.asarray(cand).reshape(ye - ys, kc)).long()
        d2 = _canonical_d2(py[:, None, :], px[cand])  # [q, kc] fp32 canonical
        # sort by (d2, index): stable sort on index first, then stable on d2
        idx_order = torch.sort(cand, dim=1, stable=True).indices
        cand = torch.gather(cand, 1, idx_order)
        d2 = torch.gather(d2, 1, idx_order)
        d_order = torch.sort(d2, dim=1, stable=True).indices
        cand = torch.gather(cand, 1, d_order)
        d2 = torch.gather(d2, 1, d_order)
        nbr[ys:ye, :kk] = cand[:, :kk] + xs
        deg[ys:ye] = kk
        if kc < n:
            # the top-k is proven iff the last candidate is strictly farther than the k-th
            unsafe = torch.nonzero(d2[:, kc - 1] <= d2[:, kk - 1]).flatten()
            for q in unsafe.tolist():
                nb, _ = knn_bruteforce(px, [0, n], py[q:q + 1], [0, 1], k)
                nbr[ys + q] = torch.where(nb[0] >= 0, nb[0] + xs, nb[0])
    return nbr, deg


def knn_graph(pos: Tensor, k: int, ptr: Sequence[int], method: str = "kdtree") -> Tensor:
    """``knn_graph(pos, k, batch=batch, loop=True)`` (pyg_randla_net.py:180):
    ``edge_index[0]`` = neighbour j (source), ``edge_index[1]`` = centre i (target),
    grouped by centre ascending, neighbours by ascending distance, self included."""
    fn = knn_kdtree if method == "kdtree" else knn_bruteforce
    nbr, deg = fn(pos, ptr, pos, ptr, k)
    mask = nbr >= 0
    centres = torch.arange(pos.shape[0]).unsqueeze(1).expand_as(nbr)
    return torch.stack([nbr[mask], centres[mask]], dim=0)


# ---------------------------------------------------------------- scatter / softmax
def scatter_sum(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_scatter ``scatter(src, index, dim=0, reduce='sum')``."""
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add(0, index, src)


def scatter_max(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_scatter ``scatter(src, index, dim=0, reduce='max')`` (value part)."""
    out = torch.full((dim_size,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)


def pyg_softmax(src: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """torch_geometric.utils.softmax(src, index) used at pyg_randla_net.py:150:
    detached per-group max, exp, group sum + 1e-16, divide."""
    src_max = scatter_max(src.detach(), index, num_nodes)
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter_sum(out, index, num_nodes) + 1e-16
    return out / out_sum.index_select(0, index)


# ------------------------------------------------------------------------ modules
class PyGBatchNorm(nn.Module):
    """torch_geometric.nn.norm.BatchNorm: wraps BatchNorm1d as ``.module``
    (state keys ``norms.N.module.*``, checkpoint Appendix C of SURVEY.md)."""

    def __init__(self, channels: int):
        super().__init__()
        self.module = nn.BatchNorm1d(channels, eps=BN_EPS, momentum=BN_MOMENTUM)

    def forward(self, x: Tensor) -> Tensor:
        return self.module(x)


class SharedMLP(nn.Module):
    """pyg_randla_net.py:97-109 = PyG ``MLP(channel_list, plain_last=False,
    act=LeakyReLU(0.2)|None, norm=batch_norm(momentum .01, eps 1e-6)|None,
    bias=..., dropout=...)``: per layer Linear -> norm -> act -> dropout."""

    def __init__(self, channels: List[int], dropout=0.0, act: bool = True, norm: bool = True, bias: bool = True):
        super().__init__()
        n = len(channels) - 1
        self.dropout = list(dropout) if isinstance(dropout, (list, tuple)) else [float(dropout)] * n
        self.act = act
        self.lins = nn.ModuleList([nn.Linear(channels[i], channels[i + 1], bias=bias) for i in range(n)])
        self.norms = nn.ModuleList([PyGBatchNorm(channels[i + 1]) if norm else nn.Identity() for i in range(n)])
        # injected dropout masks (parity harness); list aligned with layers, entries may be None
        self.injected_masks: Optional[List[Optional[Tensor]]] = None
        self.reset_parameters()

    def reset_parameters(self):
        # PyG Linear.reset_parameters == kaiming_uniform(a=sqrt(5)) for weight and
        # U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for bias, i.e. torch.nn.Linear's default.
        for lin in self.lins:
            lin.reset_parameters()

    def forward(self, x: Tensor) -> Tensor:
        for i, (lin, norm) in enumerate(zip(self.lins, self.norms)):
            x = lin(x)
            x = norm(x)
            if self.act:
                x = F.leaky_relu(x, LRELU_SLOPE)
            p = self.dropout[i]
            if self.injected_masks is not None and self.injected_masks[i] is not None and self.training:
                x = x * self.injected_masks[i]  # mask already holds 0 or 1/(1-p)
            else:
                x = F.dropout(x, p=p, training=self.training)
        return x


class LocalFeatureAggregation(nn.Module):
    """pyg_randla_net.py:112-152 (MessagePassing(aggr='add') restated)."""

    def __init__(self, channels: int):
        super().__init__()
        self.mlp_encoder = SharedMLP([10, channels // 2])
        self.mlp_attention = SharedMLP([channels, channels], bias=False, act=False, norm=False)
        self.mlp_post_attention = SharedMLP([channels, channels])

    def message(self, x_j: Tensor, pos_i: Tensor, pos_j: Tensor, index: Tensor, num_nodes: int) -> Tensor:
        pos_diff = pos_j - pos_i  # :141
        distance = torch.sqrt((pos_diff * pos_diff).sum(1, keepdim=True))  # :142
        relative_infos = torch.cat([pos_i, pos_j, pos_diff, distance], dim=1)  # :143
        local_spatial_encoding = self.mlp_encoder(relative_infos)  # :144
        local_features = torch.cat([x_j, local_spatial_encoding], dim=1)  # :145
        att_features = self.mlp_attention(local_features)  # :149
        att_scores = pyg_softmax(att_features, index, num_nodes)  # :150
        return att_scores * local_features  # :152

    def forward(self, edge_index: Tensor, x: Tensor, pos: Tensor) -> Tensor:
        j, i = edge_index[0], edge_index[1]
        msg = self.message(x.index_select(0, j), pos.index_select(0, i), pos.index_select(0, j), i, x.shape[0])
        out = scatter_sum(msg, i, x.shape[0])  # propagate(aggr="add") :122
        return self.mlp_post_attention(out)  # :123


class DilatedResidualBlock(nn.Module):
    """pyg_randla_net.py:155-189."""

    def __init__(self, num_neighbors: int, d_in: int, d_out: int, knn_method: str = "kdtree"):
        super().__init__()
        self.num_neighbors = num_neighbors
        self.knn_method = knn_method
        self.mlp1 = SharedMLP([d_in, d_out // 8])
        self.shortcut = SharedMLP([d_in, d_out], act=False)
        self.mlp2 = SharedMLP([d_out // 2, d_out], act=False)
        self.lfa1 = LocalFeatureAggregation(d_out // 4)
        self.lfa2 = LocalFeatureAggregation(d_out // 2)
        self.last_edge_index: Optional[Tensor] = None

    def forward(self, x: Tensor, pos: Tensor, ptr: Sequence[int]) -> Tensor:
        edge_index = knn_graph(pos, self.num_neighbors, ptr, self.knn_method)  # :180
        self.last_edge_index = edge_index
        shortcut_of_x = self.shortcut(x)  # :182
        x = self.mlp1(x)  # :183
        x = self.lfa1(edge_index, x, pos)  # :184
        x = self.lfa2(edge_index, x, pos)  # :185
        x = self.mlp2(x)  # :186
        return F.leaky_relu(x + shortcut_of_x, LRELU_SLOPE)  # :187


def decimation_indices(ptr: Sequence[int], decimation_factor, generator: Optional[torch.Generator] = None):
    """pyg_randla_net.py:192-231.  ``randperm`` drawn per cloud, in cloud order, from
    the torch generator (global by default, like the reference)."""
    if decimation_factor < 1:
        raise ValueError(
            "Argument `decimation_factor` should be higher than (or equal to) "
            f"1 for downsampling. (Current value: {decimation_factor})"
        )
    ptr = [int(v) for v in ptr]
    idx, new_ptr = [], [ptr[0]]
    for b in range(len(ptr) - 1):
        n = ptr[b + 1] - ptr[b]
        nd = max(1, int(n // decimation_factor))
        perm = torch.randperm(n, generator=generator) if generator is not None else torch.randperm(n)
        idx.append(ptr[b] + perm[:nd])
        new_ptr.append(new_ptr[-1] + nd)
    return torch.cat(idx, dim=0), new_ptr


def knn_interpolate(x: Tensor, pos_x: Tensor, pos_y: Tensor, ptr_x: Sequence[int], ptr_y: Sequence[int], k: int,
                    method: str = "kdtree") -> Tensor:
    """PyG ``knn_interpolate`` (pyg_randla_net.py:250 with k=1; model.py:90 with k=10)."""
    with torch.no_grad():
        fn = knn_kdtree if method == "kdtree" else knn_bruteforce
        nbr, _ = fn(pos_x, ptr_x, pos_y, ptr_y, k)
        mask = nbr >= 0
        y_idx = torch.arange(pos_y.shape[0]).unsqueeze(1).expand_as(nbr)[mask]
        x_idx = nbr[mask]
        diff = pos_x[x_idx] - pos_y[y_idx]
        squared_distance = (diff * diff).sum(dim=-1, keepdim=True)
        weights = 1.0 / torch.clamp(squared_distance, min=1e-16)
    y = scatter_sum(x[x_idx] * weights, y_idx, pos_y.shape[0])
    y = y / scatter_sum(weights, y_idx, pos_y.shape[0])
    return y


class FPModule(nn.Module):
    """pyg_randla_net.py:241-253."""

    def __init__(self, k: int, net: nn.Module, knn_method: str = "kdtree"):
        super().__init__()
        self.k = k
        self.nn = net
        self.knn_method = knn_method

    def forward(self, x, pos, ptr, x_skip, pos_skip, ptr_skip):
        x = knn_interpolate(x, pos, pos_skip, ptr, ptr_skip, self.k, self.knn_method)
        x = torch.cat([x, x_skip], dim=1)
        return self.nn(x)


class OracleRandLANet(nn.Module):
    """pyg_randla_net.py:22-88, same constructor, same state-dict keys.

    ``forward(x, pos, batch, ptr)`` as the reference; ``decimation_idx`` (list of 4
    LongTensors) may be injected so that a CUDA run and the oracle use the same
    random subsets.  The indices actually used are left in
    ``self.last_decimation_idx``.
    """

    def __init__(self, num_features: int, num_classes: int, decimation: int = 4, num_neighbors: int = 16,
                 return_logits: bool = False, knn_method: str = "kdtree"):
        super().__init__()
        self.decimation = decimation
        self.return_logits = return_logits
        d_bottleneck = max(32, num_classes, num_features)  # :40
        self.fc0 = nn.Linear(num_features, d_bottleneck)
        self.block1 = DilatedResidualBlock(num_neighbors, d_bottleneck, 32, knn_method)
        self.block2 = DilatedResidualBlock(num_neighbors, 32, 128, knn_method)
        self.block3 = DilatedResidualBlock(num_neighbors, 128, 256, knn_method)
        self.block4 = DilatedResidualBlock(num_neighbors, 256, 512, knn_method)
        self.mlp_summit = SharedMLP([512, 512])
        self.fp4 = FPModule(1, SharedMLP([512 + 256, 256]), knn_method)
        self.fp3 = FPModule(1, SharedMLP([256 + 128, 128]), knn_method)
        self.fp2 = FPModule(1, SharedMLP([128 + 32, 32]), knn_method)
        self.fp1 = FPModule(1, SharedMLP([32 + 32, d_bottleneck]), knn_method)
        self.mlp_classif = SharedMLP([d_bottleneck, 64, 32], dropout=[0.0, 0.5])
        self.fc_classif = nn.Linear(32, num_classes)
        self.last_decimation_idx: List[Tensor] = []
        self.stages: dict = {}

    def forward(self, x, pos, batch, ptr, decimation_idx: Optional[List[Tensor]] = None):
        x = x if x is not None else pos  # :56
        ptr0 = [int(v) for v in ptr]
        self.last_decimation_idx = []
        self.stages = {}

        def decimate(tensors, ptr_l, level):
            if decimation_idx is not None:
                idx = decimation_idx[level]
                _, new_ptr = _ptr_after(ptr_l, self.decimation)
            else:
                idx, new_ptr = decimation_indices(ptr_l, self.decimation)
            self.last_decimation_idx.append(idx)
            return tuple(t[idx] for t in tensors), new_ptr

        b1 = self.block1(self.fc0(x), pos, ptr0)  # :58
        self.stages["b1"] = b1
        (b1d, pos1), ptr1 = decimate((b1, pos), ptr0, 0)
        b2 = self.block2(b1d, pos1, ptr1)
        self.stages["b2"] = b2
        (b2d, pos2), ptr2 = decimate((b2, pos1), ptr1, 1)
        b3 = self.block3(b2d, pos2, ptr2)
        self.stages["b3"] = b3
        (b3d, pos3), ptr3 = decimate((b3, pos2), ptr2, 2)
        b4 = self.block4(b3d, pos3, ptr3)
        self.stages["b4"] = b4
        (b4d, pos4), ptr4 = decimate((b4, pos3), ptr3, 3)
        summit = self.mlp_summit(b4d)  # :70
        self.stages["summit"] = summit
        fp4 = self.fp4(summit, pos4, ptr4, b3d, pos3, ptr3)  # :76
        fp3 = self.fp3(fp4, pos3, ptr3, b2d, pos2, ptr2)
        fp2 = self.fp2(fp3, pos2, ptr2, b1d, pos1, ptr1)
        fp1 = self.fp1(fp2, pos1, ptr1, b1, pos, ptr0)  # :79
        self.stages["fp1"] = fp1
        h = self.mlp_classif(fp1)  # :81
        logits = self.fc_classif(h)  # :82
        if self.return_logits:
            return logits
        return logits.log_softmax(dim=-1)  # :87


def _ptr_after(ptr: Sequence[int], decimation) -> Tuple[None, List[int]]:
    new_ptr = [int(ptr[0])]
    for b in range(len(ptr) - 1):
        n = int(ptr[b + 1]) - int(ptr[b])
        new_ptr.append(new_ptr[-1] + max(1, int(n // decimation)))
    return None, new_ptr


class OracleBlock1Net(nn.Module):
    """BASELINE.json configs[0] ('1 encoder layer'): fc0 + block1 + head, no decimation,
    no decoder.  Same sub-module names as the full net."""

    def __init__(self, num_features: int, num_classes: int, num_neighbors: int = 16, knn_method: str = "kdtree"):
        super().__init__()
        d_bottleneck = max(32, num_classes, num_features)
        self.fc0 = nn.Linear(num_features, d_bottleneck)
        self.block1 = DilatedResidualBlock(num_neighbors, d_bottleneck, 32, knn_method)
        self.mlp_classif = SharedMLP([32, 64, 32], dropout=[0.0, 0.5])
        self.fc_classif = nn.Linear(32, num_classes)

    def forward(self, x, pos, batch, ptr):
        b1 = self.block1(self.fc0(x), pos, [int(v) for v in ptr])
        return self.fc_classif(self.mlp_classif(b1))


# --------------------------------------------------------------- synthetic tiles
def synthetic_tile(n: int, seed: int, num_features: int = 9, num_classes: int = 6):
    """Synthetic 50 m x 50 m Lidar-HD-like tile after the reference's transforms
    (NormalizePos /25, NullifyLowestZ, standardised intensity ...): SURVEY.md 8(d)."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.empty(n, 3)
    pos[:, 0:2] = torch.rand(n, 2, generator=g) * 2 - 1
    ground = torch.randn(n, generator=g).mul_(0.02).abs_()
    tall = (torch.rand(n, generator=g) < 0.25).float() * torch.rand(n, generator=g) * 0.6
    z = ground + tall
    pos[:, 2] = z - z.min()
    x = torch.empty(n, num_features)
    for c in range(num_features):
        if c in (0, 7):
            x[:, c] = torch.randn(n, generator=g).clamp_(-3, 3)
        elif c in (1, 2):
            x[:, c] = torch.randint(1, 8, (n,), generator=g).float() / 7
        elif c == 8:
            x[:, c] = torch.rand(n, generator=g) * 2 - 1
        else:
            x[:, c] = torch.rand(n, generator=g)
    y = torch.randint(0, num_classes, (n,), generator=g)
    return x, pos, y


def synthetic_batch(sizes: Sequence[int], seed: int = 12345, num_features: int = 9, num_classes: int = 6):
    xs, ps, ys, bs = [], [], [], []
    ptr = [0]
    for t, n in enumerate(sizes):
        x, p, y = synthetic_tile(n, seed + t, num_features, num_classes)
        xs.append(x), ps.append(p), ys.append(y)
        bs.append(torch.full((n,), t, dtype=torch.int64))
        ptr.append(ptr[-1] + n)
    return torch.cat(xs), torch.cat(ps), torch.cat(ys), torch.cat(bs), torch.tensor(ptr, dtype=torch.int64)
