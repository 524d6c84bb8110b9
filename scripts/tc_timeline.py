"""Print the in-kernel timeline (clock64 marks of one thread of CTA 0) of tc_tn_kernel for one GEMM shape."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from myria3d_b200 import _lib
from myria3d_b200.ops import _p, _stream
lib = _lib.load()
n, c1, cout = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda"
a = torch.randn(n, c1, device=dev); gy = torch.randn(n, cout, device=dev)
gw = torch.zeros(cout, c1, device=dev); gb = torch.zeros(cout, device=dev)
dbg = torch.zeros(128, dtype=torch.int64, device=dev)
lib.b200_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int64]
wsb = int(lib.b200_linear_bwd_weight_workspace_bytes(n, c1, 0, cout, 1))
ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
def run():
    _lib.check(lib.b200_linear_bwd_weight(_p(gy), _p(a), c1, c1, None, 0, 0, _p(gw), _p(gb), _p(ws), wsb, n, cout, _stream()), "x")
for _ in range(3): run()
torch.cuda.synchronize()
lib.b200_set_option(b"tc_timeline", dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.b200_set_option(b"tc_timeline", 0)
t = dbg.cpu().tolist(); k = t[127]
print(f"n={n} c1={c1} cout={cout}: kernel+reduce {e0.elapsed_time(e1)*1e3:.1f} us, {k} marks")
base = t[0]
labels = ["start", "alloc+init", "loads0 issued"]
rel = [x - base for x in t[:k]]
print("marks (cycles from start):", rel[:3], "...")
body = rel[3:k-3]
for s in range(0, min(len(body), 4 * 6), 4):
    ch = body[s:s+4]
    if len(ch) == 4: print(f"  chunk {s//4}: stage-free {ch[0]}  stored {ch[1]} (+{ch[1]-ch[0]})  next-loads-issued {ch[2]} (+{ch[2]-ch[1]})  after-barrier {ch[3]} (+{ch[3]-ch[2]})")
if len(body) >= 8:
    per = (body[-4] - body[0]) / (len(body) // 4 - 1)
    print(f"  ... {len(body)//4} chunks, {per:.0f} cycles per chunk steady state")
print("tail (last-MMA-done, epilogue-done, dealloc-done):", rel[k-3:k], " deltas:", rel[k-3]-body[-1] if body else None, rel[k-2]-rel[k-3], rel[k-1]-rel[k-2])
