"""Cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N and of where the operands live, measured
with clock64() inside the self-test kernel (one CTA, nothing else on the SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from myria3d_b200 import _lib
lib = _lib.load()
dev = "cuda"
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
for k in (64,):
    for n in (16, 32, 64, 128, 256):
        for flags, name in ((0, "A smem K / B smem K"), (2, "A smem K / B smem MN"), (3, "A smem MN / B smem MN"), (8, "A tmem / B smem K"), (10, "A tmem / B smem MN")):
            if flags & 8 and n + 3 * k // 2 > 512:
                continue
            a = torch.randn(128, k, device=dev); b = torch.randn(n, k, device=dev)
            d = torch.zeros(128, n, device=dev); st = torch.zeros(1, dtype=torch.int32, device=dev)
            lib.b200_set_option(b"tc_timeline", dbg.data_ptr())
            for _ in range(2):
                rc = lib.b200_tc_gemm_selftest(c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), c_void_p(d.data_ptr()), n, k, 6, flags,
                                               c_void_p(st.data_ptr()), c_void_p(torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
            lib.b200_set_option(b"tc_timeline", 0)
            if rc != 0:
                print(k, n, name, "rc", rc, lib.b200_last_error()); continue
            t = dbg.cpu().tolist()
            nm = 6 * k // 16
            print(f"k={k:4d} n={n:4d} {name:24s}: {nm:3d} MMAs, issue {t[0]:6d} cyc, done {t[1]:6d} cyc = {t[1] / nm:6.1f} cyc/MMA")
