"""Per-kernel micro-benchmark of libb200randla at BASELINE configs[1] shapes (16 tiles x 12 800 pts, K=16).

    python scripts/microbench.py                       # every kernel, CUDA-event timing
    python scripts/microbench.py --only lfa_bwd:16     # a subset (substring match on 'name:key')
    ncu --set full -k regex:lfa_bwd -c 1 python scripts/microbench.py --only lfa_bwd:16 --iters 1 --warmup 0

Inputs are synthetic Lidar-like tiles; neighbour tables come from the library's own kNN."""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import c_void_p
from myria3d_b200.synthetic import synthetic_batch
from myria3d_b200 import _lib, ops
from myria3d_b200.ops import _p, _stream

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--tiles", type=int, default=16)
ap.add_argument("--points", type=int, default=12800)
ap.add_argument("--json", default=None)
args = ap.parse_args()
dev = "cuda"
lib = _lib.load()
torch.manual_seed(0)

LEVELS = []  # (n_total, pts_per_tile)
n = args.points
for l in range(5):
    LEVELS.append((n * args.tiles, n))
    n = max(1, n // 4)
x0, pos0, y0, b0, ptr0 = synthetic_batch([args.points] * args.tiles, seed=12345)
pos_l = [pos0.to(dev)]
ptr_l = [ptr0.to(dev)]
for l in range(1, 5):
    sizes_prev, sizes = LEVELS[l - 1][1], LEVELS[l][1]
    idx = torch.cat([t * sizes_prev + torch.randperm(sizes_prev)[:sizes] for t in range(args.tiles)]).to(dev)
    pos_l.append(pos_l[-1][idx].contiguous())
    ptr_l.append(torch.arange(0, args.tiles + 1, device=dev, dtype=torch.int64) * sizes)

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
results = []

def bench(name, key, fn, alg_bytes=0, flops=0):
    tag = f"{name}:{key}"
    if args.only and args.only not in tag:
        return
    for _ in range(args.warmup):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    r = {"kernel": tag, "us": us, "GBps": alg_bytes / us / 1e3 if us else 0, "TFLOPs": flops / us / 1e6 if us else 0}
    results.append(r)
    print(f"{tag:42s} {us:9.1f} us   {r['GBps']:8.1f} GB/s(alg)   {r['TFLOPs']:7.2f} TFLOP/s", flush=True)

K = 16
nbr_l = []
for l in range(4):
    nt, per = LEVELS[l]
    nbr, _ = ops.knn(pos_l[l], ptr_l[l], pos_l[l], ptr_l[l], K, per, kt=16, want_dist=False)
    nbr_l.append(nbr)
    for algo in ("grid", "brute"):
        if algo == "brute" and per > 4000 and not args.only:
            continue
        bench("knn_" + algo, f"L{l}", lambda l=l, per=per, algo=algo: ops.knn(pos_l[l], ptr_l[l], pos_l[l], ptr_l[l], K, per, kt=16, want_dist=False, algo=algo),
              alg_bytes=nt * (12 + 12 + 64))
    if l < 4:
        ntc, perc = LEVELS[l + 1]
        bench("knn1_query", f"L{l + 1}->L{l}", lambda l=l, per=per, perc=perc: ops.knn(pos_l[l + 1], ptr_l[l + 1], pos_l[l], ptr_l[l], 1, per, kt=1, max_points_per_cloud=perc),
              alg_bytes=nt * (12 + 8) + ntc * 12)
    bench("edge_moments", f"L{l}", lambda l=l: ops.edge_moments(pos_l[l], nbr_l[l]), alg_bytes=nt * (12 + 64))

# LFA kernels: (level, c)
for l, c in [(0, 8), (0, 16), (1, 32), (1, 64), (2, 64), (2, 128), (3, 128), (3, 256)]:
    nt = LEVELS[l][0]
    h = c // 2
    x = torch.randn(nt, h, device=dev)
    enc_w = torch.randn(h, 7, device=dev) * 0.5
    enc_b = torch.randn(h, device=dev) * 0.1
    att_w = torch.randn(c, c, device=dev) / c ** 0.5
    att_wt = att_w.t().contiguous()
    out = torch.empty(nt, c, device=dev)
    go = torch.randn(nt, c, device=dev)
    gx, gew, geb, gaw = torch.zeros_like(x), torch.zeros_like(enc_w), torch.zeros_like(enc_b), torch.zeros_like(att_w)
    E = nt * K
    ws = torch.empty(max(16, int(lib.b200_lfa_bwd_workspace_bytes(nt, c, 16))), dtype=torch.uint8, device=dev)
    def fwd(x=x, l=l, enc_w=enc_w, enc_b=enc_b, att_wt=att_wt, out=out, nt=nt, c=c):
        _lib.check(lib.b200_lfa_fwd(_p(x), _p(pos_l[l]), _p(nbr_l[l]), _p(enc_w), _p(enc_b), _p(att_wt), _p(out), nt, c, 16, _stream()), "fwd")
    def bwd(ws=ws, x=x, l=l, enc_w=enc_w, enc_b=enc_b, att_wt=att_wt, att_w=att_w, go=go, gx=gx, gew=gew, geb=geb, gaw=gaw, nt=nt, c=c):
        _lib.check(lib.b200_lfa_bwd(_p(x), _p(pos_l[l]), _p(nbr_l[l]), _p(enc_w), _p(enc_b), _p(att_wt), _p(att_w), _p(go), _p(gx), _p(gew), _p(geb), _p(gaw), _p(ws), ws.numel(), nt, c, 16, _stream()), "bwd")
    bench("lfa_fwd", f"{c}@L{l}", fwd, alg_bytes=nt * (6 * c + 12 + 4 * K), flops=E * (2 * c * c + 2 * 7 * h))
    bench("lfa_bwd", f"{c}@L{l}", bwd, alg_bytes=nt * (8 * c + 12 + 4 * K), flops=E * (6 * c * c + 4 * 7 * h))

# per-point layers: (level, c1, c2, cout)
for l, c1, c2, cout in [(0, 9, 0, 32), (0, 32, 0, 32), (0, 32, 0, 4), (0, 16, 0, 32), (0, 32, 32, 32), (0, 32, 0, 64), (0, 64, 0, 32), (0, 16, 0, 16),
                        (1, 32, 0, 16), (1, 32, 0, 32), (1, 64, 0, 64),
                        (1, 32, 0, 128), (1, 64, 0, 128), (1, 128, 32, 32), (2, 128, 0, 256), (2, 256, 128, 128),
                        (3, 256, 0, 512), (3, 512, 256, 256), (4, 512, 0, 512)]:
    nt = LEVELS[l][0]
    a1 = torch.randn(nt, c1, device=dev)
    a2 = torch.randn(nt, c2, device=dev) if c2 else None
    w = torch.randn(cout, c1 + c2, device=dev) / (c1 + c2) ** 0.5
    bias = torch.randn(cout, device=dev)
    y = torch.empty(nt, cout, device=dev)
    gy = torch.randn(nt, cout, device=dev)
    stats = torch.zeros((int(lib.b200_linear_fwd_num_stat_partials(nt, c1, c2, cout)), 2 * cout), dtype=torch.float64, device=dev)
    ga1 = torch.empty_like(a1); ga2 = torch.empty_like(a2) if c2 else None
    gw = torch.zeros_like(w); gb = torch.zeros(cout, device=dev)
    key = f"{c1}+{c2}->{cout}@L{l}"
    fl = 2 * nt * (c1 + c2) * cout
    by = 4 * nt * (c1 + c2 + cout)
    wsi = torch.empty(max(int(lib.b200_linear_bwd_input_workspace_bytes(nt, c1, c2, cout)), 16), dtype=torch.uint8, device=dev)
    bench("linear_fwd", key, lambda a1=a1, a2=a2, w=w, bias=bias, y=y, stats=stats, nt=nt, c1=c1, c2=c2, cout=cout:
          _lib.check(lib.b200_linear_fwd(_p(a1), c1, c1, _p(a2), c2, c2, _p(w), _p(bias), _p(y), nt, cout, _p(stats), _stream()), "lf"), by, fl)
    bench("linear_bwd_input", key, lambda gy=gy, w=w, ga1=ga1, ga2=ga2, nt=nt, c1=c1, c2=c2, cout=cout:
          _lib.check(lib.b200_linear_bwd_input(_p(gy), _p(w), _p(ga1), c1, c1, _p(ga2), c2, c2, _p(wsi), wsi.numel(), nt, cout, _stream()), "lbi"), by, fl)
    wsb = int(lib.b200_linear_bwd_weight_workspace_bytes(nt, c1, c2, cout, 1))
    wsw = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    bench("linear_bwd_weight", key, lambda gy=gy, a1=a1, a2=a2, gw=gw, gb=gb, nt=nt, c1=c1, c2=c2, cout=cout, wsw=wsw, wsb=wsb:
          _lib.check(lib.b200_linear_bwd_weight(_p(gy), _p(a1), c1, c1, _p(a2), c2, c2, _p(gw), _p(gb), _p(wsw) if wsb else None, wsb, nt, cout, _stream()), "lbw"), by, fl)

if args.json:
    json.dump(results, open(args.json, "w"), indent=1)
