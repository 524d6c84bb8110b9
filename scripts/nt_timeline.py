"""In-kernel timeline (clock64 marks of warp 1 of CTA 0) of tc_nt_kernel for one Linear forward shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from myria3d_b200 import _lib
from myria3d_b200.ops import _p, _stream
lib = _lib.load()
n, k, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
a = torch.randn(n, k, device=dev); w = torch.randn(cout, k, device=dev); b = torch.randn(cout, device=dev)
y = torch.empty(n, cout, device=dev)
parts = int(lib.b200_linear_fwd_num_stat_partials(n, k, 0, cout))
stats = torch.empty(parts, 2 * cout, dtype=torch.float64, device=dev)
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
lib.b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
def run():
    _lib.check(lib.b200_linear_fwd(_p(a), k, k, None, 0, 0, _p(w), _p(b), _p(y), n, cout, _p(stats), _stream()), "x")
for _ in range(3): run()
torch.cuda.synchronize()
lib.b200_set_option(b"tc_timeline", dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.b200_set_option(b"tc_timeline", 0)
t = dbg.cpu().tolist(); cnt = t[127]
rel = [x - t[0] for x in t[:cnt]]
nch = k // 32
print(f"n={n} k={k} cout={cout}: {e0.elapsed_time(e1)*1e3:.1f} us, {cnt} marks, {nch} chunks/tile, stat partials {parts}")
print("marks: [0] start, [1] init done, then per K-chunk: F(stage free) S(stored + next loads issued) B(after barrier -> MMA issue);")
print("after a tile's chunks (from the 2nd tile on): P(previous tile's MMAs complete) E(its epilogue done)")
print("cycles:", rel[:40])
print("deltas:", [b - a for a, b in zip(rel[:39], rel[1:40])])
