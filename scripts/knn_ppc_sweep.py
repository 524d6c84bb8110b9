"""Sweep of the grid kNN's average cell occupancy (b200_set_option("knn_points_per_cell")) at the shapes of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myria3d_b200.synthetic import synthetic_batch
from myria3d_b200 import _lib, ops
lib = _lib.load()
dev = "cuda"
for per, tiles, k in ((12800, 16, 16), (3200, 16, 16), (65536, 4, 32)):
    x0, pos0, y0, b0, ptr0 = synthetic_batch([per] * tiles, seed=12345)
    pos, ptr = pos0.to(dev), ptr0.to(dev)
    ref = None
    for ppc in (3, 4, 5, 6, 8, 10, 12, 16):
        lib.b200_set_option(b"knn_points_per_cell", ppc)
        f = lambda: ops.knn(pos, ptr, pos, ptr, k, per, kt=ops.table_width(k), want_dist=False, algo="grid")
        for _ in range(3): nbr = f()[0]
        if ref is None: ref = nbr.clone()
        assert torch.equal(nbr, ref)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(10): f()
        e.record(); torch.cuda.synchronize()
        print(f"{tiles} x {per} pts, k={k}: points/cell {ppc:2d}: {s.elapsed_time(e) / 10 * 1e3:8.1f} us (whole call incl. grid build)")
lib.b200_set_option(b"knn_points_per_cell", 0)  # automatic
