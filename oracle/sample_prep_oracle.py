"""CPU oracle of the predict-time sample preparation -- TEST INFRASTRUCTURE, never imported by the
product.  Restates, line by line,

* ``get_mosaic_of_centers`` / ``split_cloud_into_samples``  myria3d/pctl/dataset/utils.py:29-38,126-158 (scipy cKDTree,
  ``query_ball_point(r=subtile_width // 2, p=inf)``; the reference yields the indices in kd-tree order, here sorted),
* ``torch_geometric.transforms.GridSampling``  (PyG 2.4: ``voxel_grid`` + ``consecutive_cluster`` + scatter mean, label =
  argmax of summed one-hots); PyG / torch_cluster are not installable here, so this is their published algorithm,
* ``torch_geometric.transforms.Center``.
Parity unpinned (the reference holds no vectors for these steps); pinned structurally by the reference's own numbers:
400 receptive fields per km^2 without overlap, 1 521 with 25 m overlap (BASELINE.json configs[3])."""
from numbers import Number
from typing import List

import numpy as np
import torch
from scipy.spatial import cKDTree


def get_mosaic_of_centers(tile_width: Number, subtile_width: Number, subtile_overlap: Number = 0):
    if subtile_overlap < 0:
        raise ValueError("datamodule.subtile_overlap must be positive.")
    xy_range = np.arange(subtile_width / 2, tile_width + (subtile_width / 2) - subtile_overlap, step=subtile_width - subtile_overlap)
    return [np.array([x, y]) for x in xy_range for y in xy_range]


def split_cloud_into_samples(pos: np.ndarray, tile_width: Number, subtile_width: Number, subtile_overlap: Number = 0) -> List[np.ndarray]:
    pos = np.asarray(pos, dtype=np.float32)
    kd_tree = cKDTree(pos[:, :2] - pos[:, :2].min(axis=0))
    out = []
    for center in get_mosaic_of_centers(tile_width, subtile_width, subtile_overlap=subtile_overlap):
        radius = subtile_width // 2  # square receptive field
        sample_idx = np.array(kd_tree.query_ball_point(center, r=radius, p=np.inf))
        if not len(sample_idx):
            continue
        out.append(np.sort(sample_idx))
    return out


def grid_sampling(pos: torch.Tensor, x: torch.Tensor, y: torch.Tensor, size: float):
    start, end = pos.min(0).values, pos.max(0).values
    c = ((pos - start) / size).long()                   # torch_cluster.grid_cluster
    num = ((end - start) / size).long() + 1
    cluster = c[:, 0] + c[:, 1] * num[0] + c[:, 2] * num[0] * num[1]
    uniq, inv = torch.unique(cluster, sorted=True, return_inverse=True)  # consecutive_cluster
    m = uniq.numel()
    cnt = torch.zeros(m, dtype=torch.float64).index_add_(0, inv, torch.ones(pos.shape[0], dtype=torch.float64))
    pos_o = (torch.zeros(m, 3, dtype=torch.float64).index_add_(0, inv, pos.double()) / cnt[:, None]).float()
    x_o = (torch.zeros(m, x.shape[1], dtype=torch.float64).index_add_(0, inv, x.double()) / cnt[:, None]).float()
    onehot = torch.nn.functional.one_hot(y)
    y_o = torch.zeros(m, onehot.shape[1], dtype=torch.int64).index_add_(0, inv, onehot).argmax(dim=-1)
    return pos_o, x_o, y_o, uniq


def center(pos: torch.Tensor) -> torch.Tensor:
    return pos - pos.mean(dim=-2, keepdim=True)


def maximum_num_nodes(n: int, num: int):
    """``MaximumNumNodes.__call__`` (myria3d/pctl/transforms/transforms.py:48-61): None when ``n <= num``, else the row
    choice ``torch.randperm(n)[:num]`` (global CPU generator, like the reference)."""
    if n <= num:
        return None
    return torch.randperm(n)[:num]


def minimum_num_nodes(n: int, num: int):
    """``MinimumNumNodes.__call__`` (transforms.py:64-81): None when ``n >= num``, else whole random permutations
    repeated ``ceil(num / n)`` times, cut at ``num``."""
    import math

    if n >= num:
        return None
    return torch.cat([torch.randperm(n) for _ in range(math.ceil(num / n))], dim=0)[:num]


def normalize_pos(pos: torch.Tensor, subtile_width: float = 50) -> torch.Tensor:
    """``NormalizePos`` (transforms.py:149-162): xy to [-1, 1] by scaling the whole cloud with 1 / (subtile_width / 2)."""
    return pos * (1 / (subtile_width / 2))
