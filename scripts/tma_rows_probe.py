"""Bring-up probe of tma_rows.cu: identity-like weights make the output a copy of chosen input columns, so a wrong
swizzle / box mapping shows up as a readable column or row permutation.  python scripts/tma_rows_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myria3d_b200 import _lib
from myria3d_b200.ops import _p, _stream

lib = _lib.load()
lib.b200_set_option(b"tma_rows", 7)
dev = "cuda"
for n, c1, c2, cout in [(8192, 32, 0, 32), (8192, 16, 0, 16), (8192, 64, 0, 64), (8192, 32, 32, 32), (8300, 32, 0, 16), (8192, 16, 0, 64),
                        (8192, 32, 0, 128), (8200, 64, 0, 128), (8192, 128, 32, 32), (8300, 128, 0, 32)]:  # the last four: weight gradient only
    k = c1 + c2
    r = torch.arange(n, device=dev, dtype=torch.float32)[:, None]
    x = r * 128 + torch.arange(k, device=dev, dtype=torch.float32)[None, :]  # exact in fp32
    a1 = x[:, :c1].contiguous()
    a2 = x[:, c1:].contiguous() if c2 else None
    w = torch.zeros(cout, k, device=dev)
    sel = [(3 * m + 1) % k for m in range(cout)]
    for m, kk in enumerate(sel):
        w[m, kk] = 1.0
    y = torch.full((n, cout), -1.0, device=dev)
    parts = int(lib.b200_linear_fwd_num_stat_partials(n, c1, c2, cout))
    stats = torch.full((parts, 2 * cout), 7.0, dtype=torch.float64, device=dev)
    rc = lib.b200_linear_fwd(_p(a1), c1, c1, _p(a2), c2, c2, _p(w), None, _p(y), n, cout, _p(stats), _stream())
    torch.cuda.synchronize()
    exp = x[:, sel]
    bad = (y != exp).nonzero()
    print(f"fwd n={n} {c1}+{c2}->{cout}: rc={rc} mismatches={bad.shape[0]} stats_sum_ok={bool(torch.allclose(stats.sum(0)[:cout], exp.double().sum(0)))}", flush=True)
    for i in range(min(6, bad.shape[0])):
        rr, cc = int(bad[i, 0]), int(bad[i, 1])
        g = float(y[rr, cc])
        print(f"   y[{rr},{cc}] = {g} (row {int(g) // 128}, col {int(g) % 128}); expected row {rr}, col {sel[cc]}")
    # weight gradient: gw[m][k] = sum_r gy[r][m] x[r][k]; gy one-hot in rows -> picks rows of x
    gy = torch.zeros(n, cout, device=dev)
    rows = [(37 * m + 5) % n for m in range(cout)]
    for m, rr in enumerate(rows):
        gy[rr, m] = 1.0
    gw = torch.zeros(cout, k, device=dev)
    gb = torch.zeros(cout, device=dev)
    rc = lib.b200_linear_bwd_weight(_p(gy), _p(a1), c1, c1, _p(a2), c2, c2, _p(gw), _p(gb), None, 0, n, cout, _stream())
    torch.cuda.synchronize()
    expw = x[rows, :]
    badw = (gw != expw).nonzero()
    print(f"tn  n={n} {c1}+{c2}->{cout}: rc={rc} mismatches={badw.shape[0]} gb_ok={bool((gb == 1).all())}", flush=True)
    for i in range(min(6, badw.shape[0])):
        mm, kk = int(badw[i, 0]), int(badw[i, 1])
        g = float(gw[mm, kk])
        print(f"   gw[{mm},{kk}] = {g} (row {int(g) // 128}, col {int(g) % 128}); expected row {rows[mm]}, col {kk}")
