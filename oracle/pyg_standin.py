"""Stand-ins for the torch_geometric / torch_cluster / torch_scatter NAMES that the reference's model file imports
(myria3d/models/modules/pyg_randla_net.py:7-19) -- TEST INFRASTRUCTURE for oracle/gen_golden_ref_model.py only.

None of those packages is installable here (no wheels, no network).  With these stand-ins the reference's OWN model code
(PyGRandLANet, SharedMLP, LocalFeatureAggregation.message, DilatedResidualBlock, decimation_indices, decimate, FPModule)
is executed unmodified; what is supplied is the published PyG 2.4 behaviour of the primitives it calls:

* ``MLP``            torch_geometric.nn.models.MLP: lins / norms ModuleLists, per layer lin -> norm -> act -> dropout,
                     ``plain_last``, ``act=None`` / ``norm=None``, list-valued dropout, ``norm_kwargs`` / ``act_kwargs``;
                     norm = the PyG BatchNorm wrapper holding ``.module = torch.nn.BatchNorm1d`` (state-dict key
                     ``norms.N.module.*``)
* ``MessagePassing`` ``propagate``: flow source_to_target (j = edge_index[0], i = edge_index[1]), ``*_i`` / ``*_j`` argument
                     collection, ``index = edge_index[1]``, aggr="add" scatter over the target nodes
* ``knn_graph``      torch_cluster: per-cloud kNN, ``loop=True`` keeps the point itself, edges grouped by centre,
                     neighbours by ascending distance  (kd-tree / brute force of oracle/randla_oracle.py)
* ``knn_interpolate``, ``softmax``, ``scatter``   as published (the oracle's restatements of the same formulas)
"""
import inspect
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

from . import randla_oracle as O


def _ptr_of(batch: torch.Tensor):
    counts = torch.bincount(batch) if batch.numel() else torch.zeros(0, dtype=torch.long)
    assert bool((batch[1:] >= batch[:-1]).all()), "clouds must be contiguous (PyG batches are)"
    return [0] + torch.cumsum(counts, 0).tolist()


class BatchNorm(nn.Module):  # torch_geometric.nn.norm.BatchNorm
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class MLP(nn.Module):  # torch_geometric.nn.models.MLP (2.4), the subset of options pyg_randla_net.py uses
    def __init__(self, channel_list, *, dropout=0.0, act="relu", act_first=False, act_kwargs=None, norm="batch_norm",
                 norm_kwargs=None, plain_last=True, bias=True):
        super().__init__()
        n = len(channel_list) - 1
        self.channel_list, self.act_first, self.plain_last = channel_list, act_first, plain_last
        self.dropout = list(dropout) if isinstance(dropout, (list, tuple)) else [float(dropout)] * n
        assert len(self.dropout) == n
        if act is None:
            self.act = None
        else:
            cls = {c.lower(): getattr(nn, c) for c in dir(nn)}[act.lower()] if isinstance(act, str) else None
            self.act = cls(**(act_kwargs or {})) if cls is not None else act
        bias = [bias] * n if isinstance(bias, bool) else bias
        self.lins = nn.ModuleList([nn.Linear(i, o, bias=b) for i, o, b in zip(channel_list[:-1], channel_list[1:], bias)])
        hidden = channel_list[1:-1] if plain_last else channel_list[1:]
        self.norms = nn.ModuleList([BatchNorm(h, **(norm_kwargs or {})) if norm is not None else nn.Identity() for h in hidden])

    def forward(self, x):
        for i, (lin, norm) in enumerate(zip(self.lins, self.norms)):
            x = lin(x)
            if self.act is not None and self.act_first:
                x = self.act(x)
            x = norm(x)
            if self.act is not None and not self.act_first:
                x = self.act(x)
            x = F.dropout(x, p=self.dropout[i], training=self.training)
        if self.plain_last:
            x = self.lins[-1](x)
            x = F.dropout(x, p=self.dropout[-1], training=self.training)
        return x


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):  # torch_scatter.scatter, sum only
    assert dim == 0 and reduce in ("sum", "add") and out is None
    return O.scatter_sum(src, index, int(dim_size if dim_size is not None else index.max() + 1))


class MessagePassing(nn.Module):  # torch_geometric.nn.conv.MessagePassing, flow="source_to_target"
    def __init__(self, aggr="add"):
        super().__init__()
        assert aggr == "add"

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        n = next(v.shape[0] for v in kwargs.values() if torch.is_tensor(v))
        args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i"):
                args[name] = kwargs[name[:-2]].index_select(0, i)
            elif name.endswith("_j"):
                args[name] = kwargs[name[:-2]].index_select(0, j)
            elif name == "index":
                args[name] = i
            else:
                args[name] = kwargs[name]
        return self.update(scatter(self.message(**args), i, dim=0, dim_size=n, reduce="sum"))

    def update(self, inputs):
        return inputs


KNN_METHOD = "kdtree"  # torch_cluster's CPU path is a kd-tree (nanoflann); "brute" is the cross-check


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    assert loop and flow == "source_to_target"
    ptr = _ptr_of(batch) if batch is not None else [0, x.shape[0]]
    return O.knn_graph(x, k, ptr, method=KNN_METHOD)


def knn_interpolate(x, pos_x, pos_y, batch_x=None, batch_y=None, k=3):
    return O.knn_interpolate(x, pos_x, pos_y, _ptr_of(batch_x), _ptr_of(batch_y), k, method=KNN_METHOD)


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    assert ptr is None and dim == 0
    return O.pyg_softmax(src, index, int(num_nodes if num_nodes is not None else index.max() + 1))


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        return _Anything(name)

    def __call__(self, *a, **k):
        return None


def modules():
    """sys.modules entries under which pyg_randla_net.py's imports resolve."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    return {
        "torch_geometric": mod("torch_geometric"),
        "torch_geometric.transforms": _Anything("torch_geometric.transforms"),
        "torch_geometric.datasets": mod("torch_geometric.datasets", ShapeNet=None),
        "torch_geometric.loader": mod("torch_geometric.loader", DataLoader=None),
        "torch_geometric.nn": mod("torch_geometric.nn", MLP=MLP),
        "torch_geometric.nn.conv": mod("torch_geometric.nn.conv", MessagePassing=MessagePassing),
        "torch_geometric.nn.pool": mod("torch_geometric.nn.pool", knn_graph=knn_graph),
        "torch_geometric.nn.unpool": mod("torch_geometric.nn.unpool", knn_interpolate=knn_interpolate),
        "torch_geometric.utils": mod("torch_geometric.utils", softmax=softmax),
        "torch_scatter": mod("torch_scatter", scatter=scatter),
        "torchmetrics": mod("torchmetrics"),
        "torchmetrics.functional": mod("torchmetrics.functional", jaccard_index=None),
    }
