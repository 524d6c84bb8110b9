"""Synthetic workload of the benchmarks: 50 m x 50 m Lidar-HD-like tiles as they leave the reference's
transforms (``NormalizePos`` x 1/25 ``myria3d/pctl/transforms/transforms.py:156-162``, ``NullifyLowestZ`` ``:141-146``,
standardised intensity / colour features ``:117-138``).  Data definition only -- no arithmetic of the hot path; kept in
the package so that ``bench.py`` and the scripts never import ``oracle/`` (which holds an identical copy for the tests;
``tests/test_cpu_host.py`` checks that the two agree bit for bit)."""
from __future__ import annotations

from typing import Sequence

import torch


def synthetic_tile(n: int, seed: int, num_features: int = 9, num_classes: int = 6):
    """``(x [n, F], pos [n, 3], y [n])``: planar position U(-1, 1), ground roughness |N(0, 0.02)| plus 25 % of the points
    on buildings / vegetation up to 0.6 (15 m), lowest z nullified; intensity-like channels N(0,1) clamped to +-3,
    return-number channels in {1..7}/7, the others uniform."""
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
    """Concatenated tiles ``(x, pos, y, batch, ptr)``; tile t uses seed ``seed + t``."""
    xs, ps, ys, bs = [], [], [], []
    ptr = [0]
    for t, n in enumerate(sizes):
        x, p, y = synthetic_tile(n, seed + t, num_features, num_classes)
        xs.append(x), ps.append(p), ys.append(y)
        bs.append(torch.full((n,), t, dtype=torch.int64))
        ptr.append(ptr[-1] + n)
    return torch.cat(xs), torch.cat(ps), torch.cat(ys), torch.cat(bs), torch.tensor(ptr, dtype=torch.int64)
