"""Shared helpers of the parity tests (oracle = checker only)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def ptr_of(sizes: Sequence[int]) -> List[int]:
    p = [0]
    for n in sizes:
        p.append(p[-1] + int(n))
    return p


def rand_cloud(sizes: Sequence[int], seed: int = 0, features: int = 9):
    g = torch.Generator().manual_seed(seed)
    n = sum(sizes)
    x = torch.rand(n, features, generator=g)
    pos = torch.rand(n, 3, generator=g)
    batch = torch.cat([torch.full((s,), i, dtype=torch.int64) for i, s in enumerate(sizes)])
    ptr = torch.tensor(ptr_of(sizes), dtype=torch.int64)
    return x, pos, batch, ptr


def assert_close(a: torch.Tensor, b: torch.Tensor, atol: float, rtol: float = 0.0, what: str = ""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    if bad.any():
        i = int(torch.argmax(err - tol))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; max |err| {float(err.max()):.3e} "
            f"(ref magnitude {float(b.abs().max()):.3e}); worst at flat index {i}: {float(a.flatten()[i])} vs {float(b.flatten()[i])}"
        )


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
