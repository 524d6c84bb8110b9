"""Predict-time sample preparation on the GPU: drop-ins for the CPU steps between a LAS tile and the
batches ``predict.py`` feeds the network,

* :func:`split_cloud_into_samples`  -- ``myria3d/pctl/dataset/utils.py:126-158`` (mosaic of receptive fields, Chebyshev
  ball query); yields the same ``idx_in_original_cloud`` sets, in ascending index order,
* :func:`grid_sampling`             -- ``torch_geometric.transforms.GridSampling(0.25)``
  (``configs/datamodule/transforms/preparations/points_budget.yaml:76-79``),
* :func:`maximum_num_nodes` / :func:`minimum_num_nodes` -- ``myria3d/pctl/transforms/transforms.py:48-84``,
* :func:`center`                    -- ``torch_geometric.transforms.Center`` (``points_budget.yaml:96-97``),
* :func:`prepare_predict_sample`    -- the ``predict`` transform list of ``points_budget.yaml:70-97`` on one sample.

Everything runs through ``libb200randla.so`` (``csrc/sample_prep.cu``); there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .ops import _call, _need_cuda, _p, _stream


def _off01(n: int, dev) -> Tensor:
    return torch.tensor([0, n], dtype=torch.int64, device=dev)


def split_cloud_into_samples(pos: Tensor, tile_width: float, subtile_width: float, subtile_overlap: float = 0.0
                             ) -> Iterator[Tensor]:
    """Indices (int64, ascending) of the points of every non-empty receptive field, in the reference's field order
    (``for x in xy_range for y in xy_range``).  ``pos`` is the fp32 ``[N, 3]`` cloud on the GPU."""
    _need_cuda(pos)
    if subtile_overlap < 0:
        raise ValueError("datamodule.subtile_overlap must be positive.")  # pctl/dataset/utils.py:30-31
    pos = pos.float().contiguous()
    n = pos.shape[0]
    g = int(_lib.load().b200_receptive_fields_per_axis(tile_width, subtile_width, subtile_overlap))
    if n == 0 or g == 0:
        return
    mn = pos[:, :2].min(dim=0).values.cpu()  # the reference subtracts pos[:, :2].min(axis=0) before building the kd-tree
    min_x, min_y = float(mn[0]), float(mn[1])
    counts = torch.empty(g * g, dtype=torch.int64, device=pos.device)
    _call("b200_receptive_fields_count", _p(pos), n, min_x, min_y, tile_width, subtile_width, subtile_overlap, _p(counts), _stream())
    offsets = torch.zeros(g * g + 1, dtype=torch.int64, device=pos.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    offsets_host = offsets.cpu().tolist()
    total = offsets_host[-1]
    if total == 0:
        return
    bufs = [torch.empty(total, dtype=torch.int32, device=pos.device) for _ in range(4)]
    cursors = torch.empty(g * g, dtype=torch.int64, device=pos.device)
    _call("b200_receptive_fields_fill", _p(pos), n, min_x, min_y, tile_width, subtile_width, subtile_overlap, _p(offsets),
          _p(cursors), _p(bufs[0]), _p(bufs[1]), _p(bufs[2]), _p(bufs[3]), _stream())
    idx = bufs[0].to(torch.int64)  # (uint32 bit patterns of indices < 2^31)
    for f in range(g * g):
        if offsets_host[f + 1] > offsets_host[f]:
            yield idx[offsets_host[f]:offsets_host[f + 1]]


def grid_sampling(pos: Tensor, x: Optional[Tensor] = None, y: Optional[Tensor] = None, size: float = 0.25
                  ) -> Tuple[Tensor, Optional[Tensor], Optional[Tensor]]:
    """``GridSampling(size)`` of one sample: one row per occupied voxel, voxels in ascending id order (x fastest, then y,
    z -- ``consecutive_cluster``), ``pos`` / ``x`` = mean of the voxel's points, ``y`` = majority label (ties: lowest)."""
    _need_cuda(pos)
    pos = pos.float().contiguous()
    n, dev = pos.shape[0], pos.device
    if n == 0:
        return pos, x, y
    xf = x.float().contiguous() if x is not None else None
    cx = xf.shape[1] if xf is not None else 0
    yl = y.to(torch.int64).contiguous() if y is not None else None
    start_end = torch.cat([pos.min(dim=0).values, pos.max(dim=0).values]).contiguous()
    key, val, kt, vt = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4))
    head = torch.empty(n, dtype=torch.int32, device=dev)
    nv = torch.zeros(1, dtype=torch.int32, device=dev)
    _call("b200_grid_sampling_sort", _p(pos), n, float(size), _p(start_end), _p(key), _p(val), _p(kt), _p(vt),
          _p(_off01(n, dev)), _p(head), _p(nv), _stream())
    run_start = torch.nonzero(head, as_tuple=False).flatten().to(torch.int32)  # ascending: one entry per voxel
    m = int(run_start.numel())
    pos_out = torch.empty(m, 3, dtype=torch.float32, device=dev)
    x_out = torch.empty(m, cx, dtype=torch.float32, device=dev) if xf is not None else None
    y_out = torch.empty(m, dtype=torch.int64, device=dev) if yl is not None else None
    ncls = int(yl.max()) + 1 if yl is not None and n else 0
    _call("b200_grid_sampling_pool", _p(key), _p(val), _p(run_start), m, n, _p(pos), _p(xf), cx, _p(yl), ncls, _p(pos_out),
          _p(x_out), _p(y_out), None, _stream())
    return pos_out, x_out, y_out


def random_permutation(n: int, device, seed: int = 0, counter: Optional[Tensor] = None, salt: int = 0) -> Tensor:
    """A uniformly random permutation of ``0..n-1`` (int64) drawn on the GPU (Philox keys + radix sort)."""
    dev = torch.device(device)
    key, perm, kt, pt = (torch.empty(max(n, 1), dtype=torch.int32, device=dev) for _ in range(4))
    _call("b200_random_permutation", n, seed & 0xFFFFFFFFFFFFFFFF, _p(counter), salt, _p(key), _p(perm), _p(kt), _p(pt),
          _p(_off01(n, dev)), _stream())
    return perm[:n].to(torch.int64)


def maximum_num_nodes(n: int, num: int, device, **rng) -> Optional[Tensor]:
    """``MaximumNumNodes(num)``: ``None`` (keep everything) when ``n <= num``, else ``randperm(n)[:num]``."""
    if n <= num:
        return None
    return random_permutation(n, device, **rng)[:num]


def minimum_num_nodes(n: int, num: int, device, seed: int = 0, counter: Optional[Tensor] = None, salt: int = 0
                      ) -> Optional[Tensor]:
    """``MinimumNumNodes(num)``: ``None`` when ``n >= num``, else ``cat([randperm(n)] * ceil(num / n))[:num]``."""
    if n >= num:
        return None
    reps = math.ceil(num / n)
    return torch.cat([random_permutation(n, device, seed=seed, counter=counter, salt=salt + 7919 * (r + 1))
                      for r in range(reps)])[:num]


def center(pos: Tensor) -> Tensor:
    """In place ``pos -= pos.mean(0)`` (fp64 accumulation)."""
    _need_cuda(pos)
    assert pos.dtype == torch.float32 and pos.is_contiguous()
    _call("b200_center_pos", _p(pos), pos.shape[0], _stream())
    return pos


def prepare_predict_sample(pos: Tensor, x: Tensor, min_nodes: int = 300, max_nodes: int = 40000, grid: float = 0.25,
                           seed: int = 0, counter: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """The ``predict`` preparation of ``points_budget.yaml:70-97`` on one receptive field (already through
    ``DropPointsByClass`` / normalisations that do not change the point set): CopyFullPos, GridSampling(0.25),
    MinimumNumNodes(300), MaximumNumNodes(40000), CopySampledPos, Center.  Returns ``x``, ``pos`` and ``copies``."""
    copies = {"pos_copy": pos.clone()}
    p, f, _ = grid_sampling(pos, x, None, grid)
    for choice in (minimum_num_nodes(p.shape[0], min_nodes, p.device, seed=seed, counter=counter, salt=1),):
        if choice is not None:
            p, f = p[choice], f[choice]
    choice = maximum_num_nodes(p.shape[0], max_nodes, p.device, seed=seed, counter=counter, salt=2)
    if choice is not None:
        p, f = p[choice], f[choice]
    p = p.contiguous()
    copies["pos_sampled_copy"] = p.clone()
    center(p)
    return {"pos": p, "x": f, "copies": copies}
