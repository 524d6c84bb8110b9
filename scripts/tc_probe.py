"""Probe how tcgen05.mma reads our shared-memory operand layout (debug aid for tc.cuh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from myria3d_b200 import _lib
lib = _lib.load()
dev = "cuda"
def run(a, b, n, k, passes=1):
    d = torch.full((128, n), float("nan"), device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ad, bd = a.to(dev), b.to(dev)
    rc = lib.b200_tc_gemm_selftest(c_void_p(ad.data_ptr()), c_void_p(bd.data_ptr()), c_void_p(d.data_ptr()), n, k, passes,
                                   c_void_p(st.data_ptr()), c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0 and int(st) == 0, (rc, int(st))
    return d.cpu()
for n, k in [(16, 8), (32, 16)]:
    b = torch.zeros(n, k)
    for i in range(n):
        for j in range(k):
            b[i, j] = 100 * i + j + 1
    print(f"=== n={n} k={k}: B[i][j] = 100 i + j + 1")
    for (r0, k0) in [(0, 0), (1, 0), (8, 0), (0, 1), (0, 4), (3, 5), (127, 7), (64, k - 1)]:
        a = torch.zeros(128, k); a[r0, k0] = 1.0
        d = run(a, b, n, k)
        rows = torch.nonzero(d.abs().sum(1) > 0).flatten().tolist()
        desc = []
        for r in rows[:4]:
            vals = d[r].tolist()
            dec = [divmod(int(round(v)) - 1, 100) if v != 0 else None for v in vals[:8]]
            desc.append(f"row {r}: " + " ".join("." if x is None else f"B[{x[0]}][{x[1]}]" for x in dec))
        print(f"A one-hot at ({r0},{k0}) -> expected row {r0} = B[0..][{k0}]; got rows {rows[:8]}{'...' if len(rows) > 8 else ''}")
        for s in desc: print("     ", s)
    # and the reverse: A encodes, B one-hot
    a = torch.zeros(128, k)
    for i in range(128):
        for j in range(k):
            a[i, j] = 10 * i + j + 1
    for (n0, k0) in [(0, 0), (1, 0), (8, 0), (0, 1), (0, 4), (5, 3)]:
        b = torch.zeros(n, k); b[n0, k0] = 1.0
        d = run(a, b, n, k)
        cols = torch.nonzero(d.abs().sum(0) > 0).flatten().tolist()
        info = []
        for c in cols[:3]:
            vals = d[:4, c].tolist()
            info.append(f"col {c}: " + " ".join(f"A[{divmod(int(round(v)) - 1, 10)[0]}][{divmod(int(round(v)) - 1, 10)[1]}]" if v != 0 else "." for v in vals))
        print(f"B one-hot at ({n0},{k0}) -> expected col {n0} = A[0..][{k0}]; got cols {cols[:8]}")
        for s in info: print("     ", s)
