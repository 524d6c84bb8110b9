"""In-kernel clock64() timeline of CTA 0 of the fused tcgen05 LFA kernels (lfa_tc.cu), per phase and tile.

    python scripts/lfa_timeline.py [c] [level]

Marks (backward): start | W staged | per tile: B1 | build | MMA1 done | E1 done | MMA3+4 done | E3 done | scatter done.
Forward: start | W staged | per tile: B1 | build | MMA1 done | E1 done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myria3d_b200.synthetic import synthetic_batch
from myria3d_b200 import _lib, ops
from myria3d_b200.ops import _p, _stream

c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = "cuda"
lib = _lib.load()
tiles, pts = 16, 12800
x0, pos0, y0, b0, ptr0 = synthetic_batch([pts] * tiles, seed=12345)
pos, per = pos0.to(dev), pts
for l in range(level):
    nxt = max(1, per // 4)
    idx = torch.cat([t * per + torch.randperm(per)[:nxt] for t in range(tiles)]).to(dev)
    pos, per = pos[idx].contiguous(), nxt
ptr = torch.arange(0, tiles + 1, device=dev, dtype=torch.int64) * per
nt = per * tiles
nbr, _ = ops.knn(pos, ptr, pos, ptr, 16, per, kt=16, want_dist=False)
h = c // 2
x = torch.randn(nt, h, device=dev)
enc_w, enc_b = torch.randn(h, 7, device=dev) * 0.5, torch.randn(h, device=dev) * 0.1
att_w = torch.randn(c, c, device=dev) / c ** 0.5
att_wt = att_w.t().contiguous()
out, go = torch.empty(nt, c, device=dev), torch.randn(nt, c, device=dev)
gx, gew, geb, gaw = torch.zeros_like(x), torch.zeros_like(enc_w), torch.zeros_like(enc_b), torch.zeros_like(att_w)
ws = torch.empty(max(256, int(lib.b200_lfa_bwd_workspace_bytes(nt, c, 16))), dtype=torch.uint8, device=dev)

def fwd():
    _lib.check(lib.b200_lfa_fwd(_p(x), _p(pos), _p(nbr), _p(enc_w), _p(enc_b), _p(att_wt), _p(out), nt, c, 16, _stream()), "fwd")
def bwd():
    _lib.check(lib.b200_lfa_bwd(_p(x), _p(pos), _p(nbr), _p(enc_w), _p(enc_b), _p(att_wt), _p(att_w), _p(go), _p(gx), _p(gew),
                                _p(geb), _p(gaw), _p(ws), ws.numel(), nt, c, 16, _stream()), "bwd")

for name, fn, per_tile in (("fwd", fwd, 4), ("bwd", bwd, 7)):
    fn(); torch.cuda.synchronize()
    dbg = torch.zeros(128, dtype=torch.int64, device=dev)
    lib.b200_set_option(b"tc_timeline", dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.b200_set_option(b"tc_timeline", 0)
    t = dbg.cpu().tolist()
    nrec = t[127]
    marks = t[:nrec]
    d = [marks[i + 1] - marks[i] for i in range(nrec - 1)]
    print(f"lfa_{name} c={c} L{level} n={nt}: {e0.elapsed_time(e1) * 1e3:.1f} us, {nrec} marks, total {marks[-1] - marks[0]} cycles")
    print("  prologue:", d[:1])
    body = d[1:]
    for i in range(0, min(len(body), per_tile * 6), per_tile):
        print("  tile", i // per_tile, body[i:i + per_tile], "sum", sum(body[i:i + per_tile]))
