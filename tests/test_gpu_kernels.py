"""Per-kernel parity tests: libb200randla (through the C ABI / ops layer) vs the CPU oracle.

Integer / index outputs are compared bit-exactly; floating outputs with the tolerance written at
each assertion.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import randla_oracle as O
from tests.helpers import assert_close, ptr_of, rand_cloud, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("sizes,k", [([700, 50, 1, 3, 1301], 16), ([5, 2100], 16), ([12, 3000, 7], 32), ([40, 900], 10),
                                     ([2500], 1)])
def test_knn_self_bit_exact(lib, sizes, k):
    from myria3d_b200 import ops

    _, pos, _, ptr = rand_cloud(sizes, seed=1)
    nbr_ref, deg_ref = O.knn_bruteforce(pos, ptr.tolist(), pos, ptr.tolist(), k)
    d2_ref = torch.full(nbr_ref.shape, float("inf"))
    m = nbr_ref >= 0
    q = torch.arange(pos.shape[0]).unsqueeze(1).expand_as(nbr_ref)
    d2_ref[m] = O._canonical_d2(pos[q[m]], pos[nbr_ref[m]])
    kt = ops.table_width(k) if k in (16, 32) else k
    nbr, d2 = ops.knn(pos.to(DEV), ptr.to(DEV), pos.to(DEV), ptr.to(DEV), k, max(sizes), kt=kt)
    assert nbr.dtype == torch.int32 and nbr.shape == (sum(sizes), kt)
    assert torch.equal(nbr.cpu()[:, :k].long(), nbr_ref), "kNN indices differ from the oracle"
    assert torch.equal(d2.cpu()[:, :k], d2_ref), "kNN squared distances differ bitwise"
    assert (nbr.cpu()[:, k:] == -1).all()
    # self is the nearest neighbour
    assert torch.equal(nbr.cpu()[:, 0].long(), torch.arange(sum(sizes)))


@pytest.mark.parametrize("k", [1, 10])
def test_knn_query_bit_exact(lib, k):
    from myria3d_b200 import ops

    sx, sy = [300, 9, 1, 1025], [1200, 36, 5, 4100]
    g = torch.Generator().manual_seed(7)
    pos_x = torch.rand(sum(sx), 3, generator=g)
    pos_y = torch.rand(sum(sy), 3, generator=g)
    nbr_ref, _ = O.knn_bruteforce(pos_x, ptr_of(sx), pos_y, ptr_of(sy), k)
    nbr, d2 = ops.knn(pos_x.to(DEV), torch.tensor(ptr_of(sx), device=DEV), pos_y.to(DEV),
                      torch.tensor(ptr_of(sy), device=DEV), k, max(sy))
    assert torch.equal(nbr.cpu().long(), nbr_ref)
    # kd-tree oracle agrees with the brute-force oracle (pins the baseline's kNN too)
    nbr_kd, _ = O.knn_kdtree(pos_x, ptr_of(sx), pos_y, ptr_of(sy), k)
    assert torch.equal(nbr_kd, nbr_ref)


def test_knn_large_property(lib):
    """BASELINE-size cloud (12 800 pts, K=16): size-independent properties + kd-tree oracle."""
    from myria3d_b200 import ops

    x, pos, y, batch, ptr = O.synthetic_batch([12800, 12800], seed=3)
    nbr, d2 = ops.knn(pos.to(DEV), ptr.to(DEV), pos.to(DEV), ptr.to(DEV), 16, 12800, kt=16)
    nbr, d2 = nbr.cpu().long(), d2.cpu()
    assert torch.equal(nbr[:, 0], torch.arange(25600))  # self first (distance 0)
    assert (d2[:, 1:] >= d2[:, :-1]).all()  # ascending
    assert ((nbr[:12800] < 12800).all() and (nbr[12800:] >= 12800).all())  # never crosses clouds
    assert (nbr.sort(dim=1).values[:, 1:] != nbr.sort(dim=1).values[:, :-1]).all()  # no duplicates
    ref, _ = O.knn_kdtree(pos, ptr.tolist(), pos, ptr.tolist(), 16)
    assert torch.equal(nbr, ref)


@pytest.mark.parametrize("k", [16, 1, 10, 32])
@pytest.mark.parametrize("dist", ["cube", "lidar", "line", "duplicates", "lattice"])
def test_knn_grid_equals_bruteforce(lib, k, dist):
    """Bucket-grid search == tiled brute force == oracle, bit for bit (indices and distances), on point
    distributions that stress the pruning: isotropic cube, flat Lidar tile, collinear points, clouds made of
    repeated points (ties decided by index), exact lattice (many equal distances)."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(hash((k, dist)) % 1000)
    sizes = [3000, 5, 1, 777, 1500]
    n = sum(sizes)
    if dist == "cube":
        pos = torch.rand(n, 3, generator=g)
    elif dist == "lidar":
        pos = O.synthetic_batch(sizes, seed=5)[1]
    elif dist == "line":
        t = torch.rand(n, 1, generator=g)
        pos = torch.cat([t * 3.0 - 1.0, t * 0.5 + 2.0, torch.zeros(n, 1)], 1)
    elif dist == "duplicates":
        base = torch.rand(40, 3, generator=g)
        pos = base[torch.randint(0, 40, (n,), generator=g)]
    else:
        ij = torch.randint(0, 24, (n, 3), generator=g).float()
        pos = ij * 0.125
    ptr = torch.tensor(ptr_of(sizes))
    pd, pt = pos.to(DEV), ptr.to(DEV)
    kt = k
    a_n, a_d = ops.knn(pd, pt, pd, pt, k, max(sizes), kt=kt, algo="brute")
    b_n, b_d = ops.knn(pd, pt, pd, pt, k, max(sizes), kt=kt, algo="grid")
    assert torch.equal(a_n, b_n), f"{int((a_n != b_n).sum())} index mismatches"
    assert torch.equal(a_d, b_d)
    ref, _ = O.knn_bruteforce(pos, ptr.tolist(), pos, ptr.tolist(), k)
    assert torch.equal(b_n.cpu().long(), ref)
    # queries != candidates, queries partly outside the candidates' bounding box
    sy = [900, 40, 3, 10, 2000]
    qpos = (torch.rand(sum(sy), 3, generator=g) * 1.6 - 0.3).to(DEV)
    qptr = torch.tensor(ptr_of(sy), device=DEV)
    c_n, c_d = ops.knn(pd, pt, qpos, qptr, k, max(sy), max_points_per_cloud=max(sizes), algo="brute")
    d_n, d_d = ops.knn(pd, pt, qpos, qptr, k, max(sy), max_points_per_cloud=max(sizes), algo="grid")
    assert torch.equal(c_n, d_n) and torch.equal(c_d, d_d)


# ---------------------------------------------------------------------------------- edge moments
def test_edge_moments(lib):
    from myria3d_b200 import ops

    sizes = [333, 7, 1200]
    _, pos, _, ptr = rand_cloud(sizes, seed=2)
    nbr, _ = O.knn_bruteforce(pos, ptr.tolist(), pos, ptr.tolist(), 16)
    m = nbr >= 0
    i = torch.arange(pos.shape[0]).unsqueeze(1).expand_as(nbr)[m]
    j = nbr[m]
    d = pos[j] - pos[i]
    dist = torch.sqrt((d * d).sum(1, keepdim=True))
    q = torch.cat([pos[i], pos[j], dist], 1).double()
    ref = torch.cat([torch.tensor([float(q.shape[0])], dtype=torch.float64), q.sum(0), (q.t() @ q).flatten()])
    out = ops.edge_moments(pos.to(DEV), nbr.int().to(DEV)).cpu()
    assert out[0].item() == q.shape[0]
    assert_close(out, ref, atol=1e-9, rtol=1e-8, what="edge moments")


@pytest.mark.parametrize("training", [True, False])
def test_encoder_fold_kernel_vs_torch_spec(lib, training):
    """b200_encoder_fold_fwd/bwd == the differentiable torch specification (randla_net.fold_encoder)."""
    import copy

    from myria3d_b200 import ops
    from myria3d_b200.randla_net import SharedMLP, fold_encoder

    g = torch.Generator().manual_seed(3)
    q = torch.rand(5000, 7, generator=g).double()
    moments = torch.cat([torch.tensor([5000.0], dtype=torch.float64), q.sum(0), (q.t() @ q).flatten()]).to(DEV)
    enc = SharedMLP([10, 32])
    bn = enc.norms[0].module
    bn.weight.data.uniform_(0.5, 1.5, generator=g), bn.bias.data.uniform_(-0.5, 0.5, generator=g)
    bn.running_mean.uniform_(-0.3, 0.3, generator=g), bn.running_var.uniform_(0.5, 1.5, generator=g)
    enc.to(DEV).train(training)
    enc2 = copy.deepcopy(enc)
    w_ref, b_ref = fold_encoder(enc, moments, 5000, training)
    w, b = ops.encoder_fold(enc2.lins[0], enc2.norms[0].module, moments, 5000, training)
    assert_close(w, w_ref, atol=1e-6, rtol=1e-5, what="enc_w")
    assert_close(b, b_ref, atol=1e-6, rtol=1e-5, what="enc_b")
    gw, gb = torch.randn(32, 7, generator=g).to(DEV), torch.randn(32, generator=g).to(DEV)
    (w_ref * gw).sum().add((b_ref * gb).sum()).backward()
    (w * gw).sum().add((b * gb).sum()).backward()
    for (n1, p1), (_, p2) in zip(enc2.named_parameters(), enc.named_parameters()):
        assert_close(p1.grad, p2.grad, atol=1e-5, rtol=1e-4, what=f"fold grad {n1}")
    for (n1, b1), (_, b2) in zip(enc2.named_buffers(), enc.named_buffers()):
        assert_close(b1, b2, atol=1e-6, rtol=1e-5, what=f"fold buffer {n1}")


# ------------------------------------------------------------------------- LFA forward / backward
@pytest.mark.parametrize("c,k", [(8, 16), (16, 16), (32, 16), (64, 16), (128, 16), (256, 16), (16, 32), (64, 32), (256, 32)])
@pytest.mark.parametrize("training", [True, False])
def test_lfa_module_parity(lib, c, k, training):
    """Product LocalFeatureAggregation (fused kernels + BN fold) vs the oracle module: output, input
    gradient and every parameter gradient.  Tolerance 2e-4 relative (fp32, different summation order)."""
    from myria3d_b200 import ops
    from myria3d_b200.randla_net import LocalFeatureAggregation, _Level

    sizes = [230, 9, 70] if k == 16 else [150, 20]
    _, pos, _, ptr = rand_cloud(sizes, seed=c + k)
    n = sum(sizes)
    g = torch.Generator().manual_seed(c)
    x = torch.randn(n, c // 2, generator=g)
    go = torch.randn(n, c, generator=g)

    ref = O.LocalFeatureAggregation(c)
    # non-trivial BatchNorm state
    for bn in (ref.mlp_encoder.norms[0].module, ref.mlp_post_attention.norms[0].module):
        bn.weight.data.uniform_(0.5, 1.5, generator=g)
        bn.bias.data.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.3, 0.3, generator=g)
        bn.running_var.uniform_(0.5, 1.5, generator=g)
    ref.train(training)
    mod = LocalFeatureAggregation(c)
    mod.load_state_dict(ref.state_dict())
    mod.to(DEV).train(training)

    edge_index = O.knn_graph(pos, k, ptr.tolist(), method="brute")
    xr = x.clone().requires_grad_(True)
    out_ref = ref(edge_index, xr, pos)
    out_ref.backward(go)

    lvl = _Level(ptr.tolist(), torch.device(DEV))
    nbr, _ = ops.knn(pos.to(DEV), lvl.ptr, pos.to(DEV), lvl.ptr, k, lvl.max_n, kt=ops.table_width(k), want_dist=False)
    moments = ops.edge_moments(pos.to(DEV), nbr) if training else None
    xg = x.to(DEV).requires_grad_(True)
    out = mod(xg, pos.to(DEV), nbr, moments, lvl.num_edges(k))
    out.backward(go.to(DEV))

    assert_close(out, out_ref, atol=2e-4 * float(out_ref.abs().max()), what=f"LFA({c}) output")
    assert rel_err(xg.grad, xr.grad) < 2e-4, f"grad_x rel err {rel_err(xg.grad, xr.grad)}"
    ref_params = dict(ref.named_parameters())
    for name, p in mod.named_parameters():
        r = ref_params[name].grad
        assert p.grad is not None, name
        e = rel_err(p.grad, r)
        scale = float(r.abs().max())
        if training and name.endswith("lins.0.bias") and "attention.lins" not in name.replace("post_attention", ""):
            # a Linear bias followed by train-mode BatchNorm has a mathematically ZERO gradient: the
            # reference's value is fp32 round-off noise, the fused path returns (almost) exact zeros
            wref = ref_params[name.replace("bias", "weight")].grad
            assert float(p.grad.abs().max()) <= 1e-3 * float(wref.abs().max()) + 1e-6, f"{name}: not ~0"
            continue
        assert e < 5e-4 or float((p.grad.cpu() - r).abs().max()) < 1e-5 * max(1.0, scale), f"{name}: rel err {e}"
    if training:
        for name, b in mod.named_buffers():
            assert_close(b, dict(ref.named_buffers())[name], atol=1e-5, rtol=1e-5, what=name)


@pytest.mark.parametrize("c,k,big", [(32, 16, False), (64, 16, False), (128, 16, False), (64, 16, True), (128, 16, True),
                                     (64, 32, False), (64, 32, True), (128, 32, False), (64, 20, False)])
def test_lfa_tensor_core_path_vs_fma_and_fp64(lib, c, k, big):
    """The tcgen05 (3xTF32, TMEM) fused LFA forward AND backward (lfa_tc.cu, the production path for c in
    {32, 64, 128}) against (a) the fp32 FMA kernels of lfa.cu (`b200_set_option("tensor_cores", 0)`) and (b) an fp64
    evaluation of pyg_randla_net.py:126-152 on the same folded encoder: pooled features, x-gradient, encoder and
    attention-weight gradients, on ragged clouds (degrees < K, partial last tile).  fp32-grade: <= 1e-5 of the
    tensor's scale against fp64."""
    from myria3d_b200 import ops
    from myria3d_b200.randla_net import _Level

    # big: > LTC_FLUSH tiles per slot (296 slots x 8 tiles x 2..4 centres), exercises the periodic dW flush
    sizes = [530, 9, 1, 77, 2500] + ([9000, 7000] if big else [])
    _, pos, _, ptr = rand_cloud(sizes, seed=c)
    n = sum(sizes)
    g = torch.Generator().manual_seed(c + 1)
    x = torch.randn(n, c // 2, generator=g)
    enc_w = torch.randn(c // 2, 7, generator=g) * 0.5
    enc_b = torch.randn(c // 2, generator=g) * 0.1
    att_w = torch.randn(c, c, generator=g) / c ** 0.5
    go = torch.randn(n, c, generator=g)
    lvl = _Level(ptr.tolist(), torch.device(DEV))
    posd = pos.to(DEV)
    kt = ops.table_width(k)
    nbr, _ = ops.knn(posd, lvl.ptr, posd, lvl.ptr, k, lvl.max_n, kt=kt, want_dist=False)

    def run():
        leaves = [t.to(DEV).requires_grad_(True) for t in (x, enc_w, enc_b, att_w)]
        out = ops.lfa_attentive_pool(leaves[0], posd, nbr, leaves[1], leaves[2], leaves[3])
        out.backward(go.to(DEV))
        torch.cuda.synchronize()
        return [out.detach().cpu()] + [t.grad.cpu() for t in leaves]

    assert lib.b200_get_option(b"tensor_cores") == 1
    tc_res = run()
    try:
        assert lib.b200_set_option(b"tensor_cores", 0) == 0
        fma_res = run()
    finally:
        lib.b200_set_option(b"tensor_cores", 1)

    # fp64 ground truth (same formulation: q = (p_i, p_j, |p_j - p_i|), folded encoder, PyG softmax)
    nb = nbr.cpu().long()
    valid = nb >= 0
    xd, wd, bd, ad = [t.double().requires_grad_(True) for t in (x, enc_w, enc_b, att_w)]
    pi = pos.double()[:, None, :].expand(-1, kt, -1)
    pj = pos.double()[nb.clamp(min=0)]
    d = pj.float() - pi.float()
    dist = torch.sqrt((d * d).sum(-1)).double()  # the kernels compute the distance in fp32 (reference order)
    q = torch.cat([pi, pj, dist[..., None]], -1)
    e = torch.nn.functional.leaky_relu(q @ wd.t() + bd, 0.2)
    f = torch.cat([xd[nb.clamp(min=0)], e], -1) * valid[..., None]
    a = f @ ad.t()
    a = a.masked_fill(~valid[..., None], float("-inf"))
    p = torch.exp(a - a.max(1, keepdim=True).values.detach())
    s = p / (p.sum(1, keepdim=True) + 1e-16)
    ref_out = (s * f).sum(1)
    ref_out.backward(go.double())
    ref = [ref_out.detach(), xd.grad, wd.grad, bd.grad, ad.grad]
    names = ["out", "grad_x", "grad_enc_w", "grad_enc_b", "grad_att_w"]
    for name, t, f_, r in zip(names, tc_res, fma_res, ref):
        scale = float(r.abs().max())
        err_tc = float((t.double() - r).abs().max()) / scale
        err_fma = float((f_.double() - r).abs().max()) / scale
        assert err_tc < 1e-5, f"{name}: tcgen05 path off by {err_tc:.2e} of scale (FMA path: {err_fma:.2e})"
        assert err_fma < 1e-5, f"{name}: FMA path off by {err_fma:.2e} of scale"


# ------------------------------------------------------------------------------ per-point layers
def test_linear_stats_cancellation(lib):
    """|mean| >> std: the fp64 statistics epilogue must still resolve the variance (2-row batches of
    tiny clouds hit this in every deep level of the [50, 50] reference test)."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(0)
    for n in (2, 300):
        a = torch.randn(n, 16, generator=g) * 1e-3 + 3.0
        w = torch.randn(8, 16, generator=g)
        b = torch.randn(8, generator=g) + 50.0
        yr = F.linear(a, w, b).double()
        y, stats = ops.linear(a.to(DEV), w.to(DEV), b.to(DEV), want_stats=True)
        stats = stats.sum(0)
        mean = stats[:8] / n
        var = stats[8:] / n - mean * mean
        assert_close(y, yr, atol=2e-5, what="y")
        ycpu = y.double().cpu()
        assert_close(var, ycpu.var(0, unbiased=False), atol=0.0, rtol=1e-6, what=f"variance n={n}")


@pytest.mark.parametrize("n,c1,c2,cout", [(5000, 128, 0, 256), (4100, 256, 128, 128), (20000, 64, 0, 64), (3000, 512, 256, 256),
                                          (1030, 64, 0, 70), (9000, 96, 0, 512), (800, 512, 0, 512), (600, 512, 256, 256)])
def test_linear_weight_grad_tensor_cores(lib, n, c1, c2, cout):
    """Weight / bias gradients through the tcgen05 3xTF32 split-K kernel (tc_gemm.cu) vs fp64: fp32-grade."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n)
    a1 = torch.randn(n, c1, generator=g)
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, generator=g)
    inp = torch.cat([a1, a2], 1) if c2 else a1
    gw_ref = gy.double().t() @ inp.double()
    gb_ref = gy.double().sum(0)
    ag = [t.to(DEV).requires_grad_(True) for t in (a1, w, b)]
    a2g = a2.to(DEV) if c2 else None
    y = ops.linear(ag[0], ag[1], ag[2], a2=a2g)
    y.backward(gy.to(DEV))
    assert rel_err(ag[1].grad, gw_ref) < 5e-6, rel_err(ag[1].grad, gw_ref)  # fp32 accumulation over n rows
    assert rel_err(ag[2].grad, gb_ref) < 5e-6, rel_err(ag[2].grad, gb_ref)


@pytest.mark.parametrize("n,c1,c2,cout", [(204800, 32, 0, 32), (60001, 9, 0, 32), (8192, 32, 32, 32), (20000, 17, 0, 6),
                                          (51200, 32, 0, 4), (30000, 64, 0, 32), (10000, 8, 0, 8), (9000, 16, 0, 16),
                                          (12345, 24, 8, 20), (8200, 16, 48, 5), (40000, 32, 0, 33)])
def test_linear_row_streaming_kernel(lib, n, c1, c2, cout):
    """linear_rows.cu (one thread per row; levels 0-1, <= 64 channels on both sides): output, fp64 BatchNorm column
    statistics (with a large common offset: var = E[y^2] - E[y]^2 must not cancel), input gradients of both segments,
    vs fp64; vectorised and scalar load / store paths, ragged last tile, padded K and cout."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + cout)
    a1 = torch.randn(n, c1, generator=g)
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g) + 30.0  # |mean| >> std
    gy = torch.randn(n, cout, generator=g)
    inp = (torch.cat([a1, a2], 1) if c2 else a1).double()
    y_ref = inp @ w.double().t() + b.double()
    ga_ref = gy.double() @ w.double()
    ag = [t.to(DEV).requires_grad_(True) for t in (a1, w, b)] + ([a2.to(DEV).requires_grad_(True)] if c2 else [])
    y, stats = ops.linear(ag[0], ag[1], ag[2], a2=ag[3] if c2 else None, want_stats=True)
    y.backward(gy.to(DEV))
    assert_close(y, y_ref, atol=2e-5, rtol=2e-6, what="linear_rows y")
    st = stats.sum(0)
    ycpu = y.detach().double().cpu()
    mean = st[:cout] / n
    var = st[cout:] / n - mean * mean
    assert_close(mean, ycpu.mean(0), atol=0.0, rtol=1e-9, what="column means")
    assert_close(var, ycpu.var(0, unbiased=False), atol=0.0, rtol=1e-6, what="column variances")
    assert rel_err(ag[0].grad, ga_ref[:, :c1]) < 2e-6, rel_err(ag[0].grad, ga_ref[:, :c1])
    if c2:
        assert rel_err(ag[3].grad, ga_ref[:, c1:]) < 2e-6
    # only one of the two input gradients wanted (the other output pointer is null)
    if c2:
        a1n, a2g = a1.to(DEV), a2.to(DEV).requires_grad_(True)
        ops.linear(a1n, ag[1].detach(), ag[2].detach(), a2=a2g).backward(gy.to(DEV))
        assert rel_err(a2g.grad, ga_ref[:, c1:]) < 2e-6


@pytest.mark.parametrize("n,c1,c2,cout", [(204800, 32, 0, 32), (51200, 16, 0, 32), (8269, 32, 32, 32), (30000, 64, 0, 32),
                                          (20000, 32, 0, 64), (9000, 16, 0, 16), (12345, 64, 0, 64), (10000, 32, 0, 16),
                                          (8200, 16, 0, 64), (40001, 64, 0, 16), (204800, 32, 32, 32),
                                          (51200, 32, 0, 128), (20001, 64, 0, 128), (51200, 128, 32, 32), (12800, 128, 0, 32)])
def test_linear_tma_rows(lib, n, c1, c2, cout):
    """tma_rows.cu (2-D tensor-map TMA boxes through an mbarrier ring, persistent CTAs; 16/32/64-channel layers on
    >= 8192 rows): forward + fp64 BatchNorm column statistics (|mean| >> std), input gradients of both segments, weight
    and bias gradients vs fp64, ragged last tile (rows past n are zero-filled by the TMA unit), and agreement with the
    register-staged kernels the option bits replace."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + cout)
    a1 = torch.randn(n, c1, generator=g) + 0.2
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g) + 30.0  # |mean| >> std
    gy = torch.randn(n, cout, generator=g)
    inp = (torch.cat([a1, a2], 1) if c2 else a1).double()
    y_ref = inp @ w.double().t() + b.double()
    ga_ref = gy.double() @ w.double()
    gw_ref = gy.double().t() @ inp
    gb_ref = gy.double().sum(0)

    def run():
        ag = [t.to(DEV).requires_grad_(True) for t in (a1, w, b)] + ([a2.to(DEV).requires_grad_(True)] if c2 else [])
        y, stats = ops.linear(ag[0], ag[1], ag[2], a2=ag[3] if c2 else None, want_stats=True)
        y.backward(gy.to(DEV))
        torch.cuda.synchronize()
        return y.detach(), stats, ag

    before = int(lib.b200_get_option(b"tma_rows"))
    try:
        lib.b200_set_option(b"tma_rows", 7)
        y, stats, ag = run()
        lib.b200_set_option(b"tma_rows", 0)
        y0, stats0, ag0 = run()
    finally:
        lib.b200_set_option(b"tma_rows", before)
    wide = (c1 + c2 == 64 and cout == 64) or c1 + c2 > 64 or cout > 64
    assert_close(y, y_ref, atol=3e-5 if wide else 2e-5, rtol=1e-5 if wide else 2e-6, what="tma_rows y")
    st = stats.sum(0)
    ycpu = y.double().cpu()
    mean = st[:cout] / n
    var = st[cout:] / n - mean * mean
    # (64 x 64 and the >= 128-wide shapes: forward / input gradient stay on tc_nt.cu / the tile FMA kernels -- statistics of
    # the tcgen05 epilogue, 3xTF32 products; their weight gradients still go through tma_rows_tn)
    wide = (c1 + c2 == 64 and cout == 64) or c1 + c2 > 64 or cout > 64
    assert_close(mean, ycpu.mean(0), atol=0.0, rtol=1e-7 if wide else 1e-9, what="column means")
    assert_close(var, ycpu.var(0, unbiased=False), atol=0.0, rtol=1e-4 if wide else 1e-6, what="column variances")
    assert rel_err(ag[0].grad, ga_ref[:, :c1]) < (1e-5 if wide else 2e-6), rel_err(ag[0].grad, ga_ref[:, :c1])
    if c2:
        assert rel_err(ag[3].grad, ga_ref[:, c1:]) < (1e-5 if wide else 2e-6), rel_err(ag[3].grad, ga_ref[:, c1:])
    assert rel_err(ag[1].grad, gw_ref) < 5e-6, rel_err(ag[1].grad, gw_ref)  # fp32 accumulation over n rows
    assert rel_err(ag[2].grad, gb_ref) < 5e-6, rel_err(ag[2].grad, gb_ref)
    # the kernels they stand in for agree to fp32 round-off (different summation orders)
    assert rel_err(y, y0.double()) < 1e-5 and rel_err(ag[0].grad, ag0[0].grad.double()) < 1e-5
    assert rel_err(ag[1].grad, ag0[1].grad.double()) < 1e-5
    # one partial row per CTA, the remaining rows of the caller's buffer are exact zeros
    assert stats.shape == stats0.shape
    if n >= 204800:
        assert float(stats[2 * 148:].abs().max()) == 0.0, "the TMA kernel did not run (partials beyond the grid are not zero)"


def test_linear_tma_rows_kernels_run(lib):
    """The option bits really select the TMA kernels (kernel names from CUPTI) and a null input-gradient pointer of
    one segment is honoured."""
    from myria3d_b200 import ops
    from torch.profiler import ProfilerActivity, profile

    g = torch.Generator().manual_seed(5)
    n = 16384
    a1, a2 = torch.randn(n, 32, generator=g).to(DEV), torch.randn(n, 32, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(32, 64, generator=g) / 8).to(DEV).requires_grad_(True)
    b = torch.randn(32, generator=g).to(DEV).requires_grad_(True)
    gy = torch.randn(n, 32, generator=g).to(DEV)
    before = int(lib.b200_get_option(b"tma_rows"))
    try:
        lib.b200_set_option(b"tma_rows", 7)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            ops.linear(a1, w, b, a2=a2).backward(gy)
            torch.cuda.synchronize()
    finally:
        lib.b200_set_option(b"tma_rows", before)
    names = " ".join(e.key for e in prof.key_averages())
    assert names.count("tma_rows_nn_kernel") >= 1 and "tma_rows_tn_kernel" in names, names
    ga_ref = gy.double() @ w.detach().double()
    assert rel_err(a2.grad, ga_ref[:, 32:]) < 2e-6


@pytest.mark.parametrize("n,c1,c2,cout,bias", [(204800, 32, 0, 4, True), (204800, 9, 0, 32, True), (204800, 32, 0, 33, True),
                                               (51200, 32, 0, 4, True), (60001, 17, 0, 32, True), (4096, 64, 0, 12, False),
                                               (204800, 32, 0, 64, True), (7777, 24, 8, 20, True)])
def test_linear_weight_grad_narrow_tensor_cores(lib, n, c1, c2, cout, bias):
    """Weight / bias gradients of the narrow level-0/1 layers through tc_skinny.cu (tcgen05 kind::f16 on bf16 x 3
    operands read MN-major from their row-major layout, K = rows) vs fp64 and vs the FMA fallback: fp32-grade."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + cout)
    a1 = torch.randn(n, c1, generator=g) + 0.3
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g) if bias else None
    gy = torch.randn(n, cout, generator=g)
    inp = torch.cat([a1, a2], 1) if c2 else a1
    gw_ref = gy.double().t() @ inp.double()
    gb_ref = gy.double().sum(0)

    def run():
        ag = [t.to(DEV).requires_grad_(True) for t in ((a1, w, b) if bias else (a1, w))]
        y = ops.linear(ag[0], ag[1], ag[2] if bias else None, a2=a2.to(DEV) if c2 else None)
        y.backward(gy.to(DEV))
        return ag[1].grad, (ag[2].grad if bias else None)

    gw_tc, gb_tc = run()
    try:
        lib.b200_set_option(b"tensor_cores", 0)
        gw_fma, gb_fma = run()
    finally:
        lib.b200_set_option(b"tensor_cores", 1)
    assert rel_err(gw_tc, gw_ref) < 5e-6, (rel_err(gw_tc, gw_ref), rel_err(gw_fma, gw_ref))
    assert rel_err(gw_fma, gw_ref) < 5e-6
    if bias:
        assert rel_err(gb_tc, gb_ref) < 5e-6 and rel_err(gb_fma, gb_ref) < 5e-6, (rel_err(gb_tc, gb_ref), rel_err(gb_fma, gb_ref))


@pytest.mark.parametrize("n,c1,c2,cout", [(5000, 128, 0, 256), (4100, 256, 128, 128), (30000, 64, 0, 64), (3000, 512, 256, 256),
                                          (1030, 64, 0, 70), (1024, 96, 32, 512), (2500, 64, 64, 100), (40000, 32, 0, 128),
                                          (800, 512, 0, 512), (700, 512, 256, 256), (513, 64, 0, 64)])
def test_linear_tensor_core_forward_and_input_grad(lib, n, c1, c2, cout):
    """Layers with >= 64 input and output channels run on tcgen05 (tc_nt.cu, 3xTF32, channels on TMEM lanes): output,
    BatchNorm column statistics and input gradients against fp64 -- fp32-grade (rel 2e-6), ragged last row tile,
    channel counts that are not multiples of 128, two-segment input/output."""
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + cout)
    a1 = torch.randn(n, c1, generator=g)
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, generator=g)
    inp = (torch.cat([a1, a2], 1) if c2 else a1).double()
    y_ref = inp @ w.double().t() + b.double()
    ga_ref = gy.double() @ w.double()
    ag = [t.to(DEV).requires_grad_(True) for t in (a1, w, b)] + ([a2.to(DEV).requires_grad_(True)] if c2 else [])
    y, stats = ops.linear(ag[0], ag[1], ag[2], a2=ag[3] if c2 else None, want_stats=True)
    y.backward(gy.to(DEV))
    assert rel_err(y, y_ref) < 1e-5, rel_err(y, y_ref)  # 3xTF32, tensor-core accumulation over up to 768 terms (fp32 FMA: ~2e-6)
    assert_close(y, y_ref, atol=3e-6 * (c1 + c2) ** 0.5, rtol=1e-5, what="linear y (tcgen05)")  # K-term sums, |y| up to ~7
    stats = stats.sum(0)
    assert_close(stats[:cout], y_ref.sum(0), atol=1e-3, rtol=1e-5, what="column sums")
    assert_close(stats[cout:], (y_ref ** 2).sum(0), atol=1e-3, rtol=1e-5, what="column sums of squares")
    assert rel_err(ag[0].grad, ga_ref[:, :c1]) < 1e-5, rel_err(ag[0].grad, ga_ref[:, :c1])
    if c2:
        assert rel_err(ag[3].grad, ga_ref[:, c1:]) < 1e-5, rel_err(ag[3].grad, ga_ref[:, c1:])
    # and bit-for-bit insensitive to what lies beyond the last row (no stale shared memory / TMEM in the ragged tile)
    y2 = ops.linear(ag[0].detach(), ag[1].detach(), ag[2].detach(), a2=ag[3].detach() if c2 else None)
    assert torch.equal(y2, y.detach())


@pytest.mark.parametrize("n,c1,c2,cout", [(1000, 9, 0, 32), (777, 32, 32, 32), (130, 512, 256, 256), (2048, 32, 0, 7),
                                          (65, 64, 0, 64), (3, 4, 0, 8), (515, 128, 32, 32)])
def test_linear_fwd_bwd(lib, n, c1, c2, cout):
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n)
    a1 = torch.randn(n, c1, generator=g)
    a2 = torch.randn(n, c2, generator=g) if c2 else None
    w = torch.randn(cout, c1 + c2, generator=g) / (c1 + c2) ** 0.5
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, generator=g)
    ar = [t.clone().requires_grad_(True) for t in (a1, w, b)] + ([a2.clone().requires_grad_(True)] if c2 else [])
    inp = torch.cat([ar[0], ar[3]], 1) if c2 else ar[0]
    yr = F.linear(inp, ar[1], ar[2])
    yr.backward(gy)

    ag = [t.to(DEV).requires_grad_(True) for t in (a1, w, b)] + ([a2.to(DEV).requires_grad_(True)] if c2 else [])
    y, stats = ops.linear(ag[0], ag[1], ag[2], a2=ag[3] if c2 else None, want_stats=True)
    y.backward(gy.to(DEV))
    tol = 1e-5 * (c1 + c2) ** 0.5
    assert_close(y, yr, atol=tol, rtol=1e-5, what="linear y")
    stats = stats.sum(0)  # per-row-tile partials
    assert_close(stats[:cout], yr.double().sum(0), atol=1e-3, rtol=1e-5, what="column sums")
    assert_close(stats[cout:], (yr.double() ** 2).sum(0), atol=1e-3, rtol=1e-5, what="column sums of squares")
    assert_close(ag[0].grad, ar[0].grad, atol=1e-5 * cout ** 0.5 * 3, rtol=1e-5, what="grad a1")
    if c2:
        assert_close(ag[3].grad, ar[3].grad, atol=1e-5 * cout ** 0.5 * 3, rtol=1e-5, what="grad a2")
    assert rel_err(ag[1].grad, ar[1].grad) < 1e-5, "grad w"
    assert rel_err(ag[2].grad, ar[2].grad) < 1e-5, "grad b"


@pytest.mark.parametrize("n,c", [(1000, 32), (257, 512), (5000, 4), (2, 64), (300, 40)])
@pytest.mark.parametrize("dual", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act(lib, n, c, dual, training):
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + c)
    y1 = torch.randn(n, c, generator=g) * 2 + 0.7
    y2 = torch.randn(n, c, generator=g) - 0.3
    go = torch.randn(n, c, generator=g)

    def make_bn():
        bn = torch.nn.BatchNorm1d(c, eps=1e-6, momentum=0.01)
        bn.weight.data.uniform_(0.5, 1.5, generator=g)
        bn.bias.data.uniform_(-0.5, 0.5, generator=g)
        bn.running_mean.uniform_(-0.3, 0.3, generator=g)
        bn.running_var.uniform_(0.5, 1.5, generator=g)
        return bn

    import copy

    bn1, bn2 = make_bn(), make_bn()
    gb1, gb2 = copy.deepcopy(bn1).to(DEV), copy.deepcopy(bn2).to(DEV)
    for m in (bn1, bn2, gb1, gb2):
        m.train(training)
    r1, r2 = y1.clone().requires_grad_(True), y2.clone().requires_grad_(True)
    pre = bn1(r1) + (bn2(r2) if dual else 0)
    out_ref = F.leaky_relu(pre, 0.2)
    out_ref.backward(go)

    d1, d2 = y1.to(DEV).requires_grad_(True), y2.to(DEV).requires_grad_(True)

    def stats(t):
        td = t.detach().double()
        return torch.cat([td.sum(0), (td * td).sum(0)])[None, :].contiguous() if training else None

    out = ops.bn_act(d1, stats(d1), gb1, 0.2, y2=d2 if dual else None, stats2=stats(d2) if dual else None,
                     bn2=gb2 if dual else None)
    out.backward(go.to(DEV))
    assert_close(out, out_ref, atol=2e-5, rtol=1e-5, what="bn_act out")
    assert_close(d1.grad, r1.grad, atol=2e-5, rtol=1e-4, what="grad y1")
    assert rel_err(gb1.weight.grad, bn1.weight.grad) < 1e-4
    assert rel_err(gb1.bias.grad, bn1.bias.grad) < 1e-4
    assert_close(gb1.running_mean, bn1.running_mean, atol=1e-6, rtol=1e-5, what="running_mean")
    assert_close(gb1.running_var, bn1.running_var, atol=1e-6, rtol=1e-5, what="running_var")
    assert int(gb1.num_batches_tracked) == int(bn1.num_batches_tracked)
    if dual:
        assert_close(d2.grad, r2.grad, atol=2e-5, rtol=1e-4, what="grad y2")
        assert rel_err(gb2.weight.grad, bn2.weight.grad) < 1e-4
        assert_close(gb2.running_var, bn2.running_var, atol=1e-6, rtol=1e-5, what="running_var 2")


@pytest.mark.parametrize("n,c", [(12800, 128), (51200, 32), (3200, 512), (800, 512), (12345, 64), (204800, 8), (4099, 20)])
@pytest.mark.parametrize("dual", [False, True])
def test_bn_backward_one_launch_matches_two_kernels(lib, n, c, dual):
    """b200_affine_act_bwd can run reduce + apply of the small levels as ONE kernel with a grid barrier (n * c <= 2^21;
    option bn_backward_fused, off by default because it measured slower):
    against fp64 torch autograd, against the two-kernel path (option bn_backward_fused = 0) to fp32 round-off, repeated
    calls (the barrier counter and the sums come from the zeroed scratch arena every time), and the kernel name from CUPTI."""
    import copy

    from myria3d_b200 import ops
    from torch.profiler import ProfilerActivity, profile

    g = torch.Generator().manual_seed(n + c)
    y1 = (torch.randn(n, c, generator=g) * 2 + 0.7).to(DEV)
    y2 = (torch.randn(n, c, generator=g) - 0.3).to(DEV)
    go = torch.randn(n, c, generator=g).to(DEV)
    bn1 = torch.nn.BatchNorm1d(c, eps=1e-6, momentum=0.01)
    bn1.weight.data.uniform_(0.5, 1.5, generator=g)
    bn1.bias.data.uniform_(-0.5, 0.5, generator=g)
    bn2 = copy.deepcopy(bn1)
    bn1, bn2 = bn1.to(DEV).train(), bn2.to(DEV).train()

    def stats(t):
        td = t.detach().double()
        return torch.cat([td.sum(0), (td * td).sum(0)])[None, :].contiguous()

    def run():
        m1, m2 = copy.deepcopy(bn1), copy.deepcopy(bn2)
        d1, d2 = y1.clone().requires_grad_(True), y2.clone().requires_grad_(True)
        out = ops.bn_act(d1, stats(d1), m1, 0.2, y2=d2 if dual else None, stats2=stats(d2) if dual else None, bn2=m2 if dual else None)
        out.backward(go)
        torch.cuda.synchronize()
        return d1.grad, (d2.grad if dual else None), m1.weight.grad, m1.bias.grad

    before = int(lib.b200_get_option(b"bn_backward_fused"))
    try:
        lib.b200_set_option(b"bn_backward_fused", 1)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            a = run()
        a2 = run()
        lib.b200_set_option(b"bn_backward_fused", 0)
        b = run()
    finally:
        lib.b200_set_option(b"bn_backward_fused", before)
    names = " ".join(e.key for e in prof.key_averages())
    small = n * c <= (1 << 21) and c % 4 == 0
    assert ("affine_act_bwd_fused_kernel" in names) == small, names
    # fp64 reference through torch autograd
    r1, r2 = y1.double().clone().requires_grad_(True), y2.double().clone().requires_grad_(True)
    q1, q2 = copy.deepcopy(bn1).double(), copy.deepcopy(bn2).double()
    pre = q1(r1) + (q2(r2) if dual else 0)
    F.leaky_relu(pre, 0.2).backward(go.double())
    assert rel_err(a[0], r1.grad) < 2e-5, rel_err(a[0], r1.grad)
    assert rel_err(a[2], q1.weight.grad) < 1e-5 and rel_err(a[3], q1.bias.grad) < 1e-5
    if dual:
        assert rel_err(a[1], r2.grad) < 2e-5
    for x, y in zip(a, b):  # one launch == two launches (fp64 atomics in a different order: round-off of the fp32 results)
        if x is not None:
            assert rel_err(x, y.double()) < 1e-6, rel_err(x, y.double())
    for x, y in zip(a, a2):
        if x is not None:
            assert rel_err(x, y.double()) < 1e-6


def test_bn_single_row_raises(lib):
    from myria3d_b200 import ops

    bn = torch.nn.BatchNorm1d(8).to(DEV)
    y = torch.randn(1, 8, device=DEV)
    with pytest.raises(ValueError):
        ops.bn_act(y, torch.zeros((1, 16), dtype=torch.float64, device=DEV), bn, 0.2)


# ------------------------------------------------------------------------------- index / scatter
@pytest.mark.parametrize("c", [3, 32, 7, 512])
def test_gather_rows_bit_exact(lib, c):
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(c)
    x = torch.randn(1000, c, generator=g)
    idx = torch.randperm(1000, generator=g)[:333]
    xg = x.to(DEV).requires_grad_(True)
    out = ops.gather_rows(xg, idx.to(DEV))
    assert torch.equal(out.cpu(), x[idx])
    go = torch.randn(333, c, generator=g)
    out.backward(go.to(DEV))
    ref = torch.zeros_like(x).index_add(0, idx, go)
    assert torch.equal(xg.grad.cpu(), ref)


@pytest.mark.parametrize("k,c", [(1, 512), (1, 32), (10, 7), (3, 6)])
def test_knn_interpolate_bit_exact(lib, k, c):
    """Forward bit-exact vs the oracle's restatement of PyG knn_interpolate ((x*w)/w rounding included)."""
    from myria3d_b200 import ops

    sx, sy = [50, 3, 200], [200, 12, 800]
    g = torch.Generator().manual_seed(k * 100 + c)
    pos_x = torch.rand(sum(sx), 3, generator=g)
    pos_y = torch.rand(sum(sy), 3, generator=g)
    pos_y[5] = pos_x[7]  # an exact hit: d2 = 0 -> w = 1e16
    x = torch.randn(sum(sx), c, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = O.knn_interpolate(xr, pos_x, pos_y, ptr_of(sx), ptr_of(sy), k, method="brute")
    go = torch.randn(sum(sy), c, generator=g)
    ref.backward(go)

    nbr, d2 = ops.knn(pos_x.to(DEV), torch.tensor(ptr_of(sx), device=DEV), pos_y.to(DEV),
                      torch.tensor(ptr_of(sy), device=DEV), k, max(sy))
    xg = x.to(DEV).requires_grad_(True)
    out = ops.knn_interpolate_from_table(xg, nbr, d2, k)
    out.backward(go.to(DEV))
    assert torch.equal(out.cpu(), ref.detach()), f"max diff {(out.cpu() - ref.detach()).abs().max()}"
    assert_close(xg.grad, xr.grad, atol=1e-5, rtol=1e-5, what="interp grad")


# ------------------------------------------------------------------------------ decimation draw
def test_decimation_draw_kernel(lib):
    """b200_decimation_draw == the distribution of ``ptr[b] + torch.randperm(n_b)[:max(1, n_b // 4)]``
    (pyg_randla_net.py:216-221): valid (distinct, in-cloud, right counts), deterministic in (seed, counter, salt),
    fresh per counter value, and uniform: chi-square of the inclusion counts and of the FIRST drawn position."""
    from myria3d_b200 import ops
    from myria3d_b200.randla_net import decimation_sizes

    sizes = [12800, 1, 3, 50, 4097, 65536]
    ptr_h = ptr_of(sizes)
    new_h = decimation_sizes(ptr_h, 4)
    ptr, new_ptr = torch.tensor(ptr_h, device=DEV), torch.tensor(new_h, device=DEV)
    kept = [new_h[i + 1] - new_h[i] for i in range(len(sizes))]
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    idx = ops.decimation_draw(ptr, new_ptr, max(kept), new_h[-1], 1234, counter, 2).cpu()
    for b, (n, k) in enumerate(zip(sizes, kept)):
        part = idx[new_h[b]:new_h[b + 1]]
        assert k == max(1, n // 4) and part.numel() == k
        assert int(part.min()) >= ptr_h[b] and int(part.max()) < ptr_h[b + 1]
        assert part.unique().numel() == k, "duplicates in the draw"
    again = ops.decimation_draw(ptr, new_ptr, max(kept), new_h[-1], 1234, counter, 2).cpu()
    assert torch.equal(idx, again), "not a function of (seed, counter, salt)"
    other_salt = ops.decimation_draw(ptr, new_ptr, max(kept), new_h[-1], 1234, counter, 3).cpu()
    assert not torch.equal(idx, other_salt)
    ops.counter_add(counter, 1)
    fresh = ops.decimation_draw(ptr, new_ptr, max(kept), new_h[-1], 1234, counter, 2).cpu()
    assert int(counter) == 1 and not torch.equal(idx, fresh)
    assert not torch.equal(idx[:3200].sort().values, fresh[:3200].sort().values), "same subset twice"

    # uniformity on a small cloud: 64 points, 16 kept, 4000 draws
    ptr2, new2 = torch.tensor([0, 64], device=DEV), torch.tensor([0, 16], device=DEV)
    trials = 4000
    incl = torch.zeros(64)
    first = torch.zeros(64)
    pair_order = 0
    for t in range(trials):
        d = ops.decimation_draw(ptr2, new2, 16, 16, 99, counter, 0).cpu()
        ops.counter_add(counter, 1)
        incl[d] += 1
        first[d[0]] += 1
        pair_order += int(d[0] < d[1])
    exp_incl, exp_first = trials * 16 / 64, trials / 64
    chi_incl = float(((incl - exp_incl) ** 2 / (exp_incl * (1 - 16 / 64))).sum())  # ~ chi2(63)
    chi_first = float(((first - exp_first) ** 2 / exp_first).sum())              # ~ chi2(63)
    assert chi_incl < 120 and chi_first < 120, (chi_incl, chi_first)              # p ~ 1e-5 at 63 dof
    assert abs(pair_order / trials - 0.5) < 0.04, "the order inside the subset is not random"


# ------------------------------------------------------------------------------ tcgen05 building blocks
@pytest.mark.parametrize("flags", [0, 1, 2, 3, 4, 7, 8, 14])
@pytest.mark.parametrize("n,k", [(128, 64), (64, 32), (256, 64), (16, 16), (128, 16), (32, 128)])
def test_tcgen05_gemm_selftest(lib, n, k, flags):
    """tcgen05.mma through our shared-memory descriptors + TMEM load/store path vs an fp64 product: plain TF32 ~1e-3
    relative, 3xTF32 and bf16 x 3 (kind::f16, six cross products) ~1e-6 (fp32-grade).  flags bit 0/1: operand A/B
    staged transposed and read through the MN-major descriptor (how the fused LFA backward re-reads dA, F and W_att in
    place; 16-bit operands only); bit 2: the accumulator is pre-initialised with tcgen05.st; bit 3: A lives in tensor
    memory (tcgen05.mma with a TMEM A operand), as W_att does in the fused LFA kernels.

    The tensor core's fp32 accumulate truncates (~6e-8 relative per accumulation step, always towards zero), so the
    error of the split products grows with the number of MMAs chained on one accumulator: 5e-6 covers 48 steps."""
    from ctypes import c_void_p

    g = torch.Generator().manual_seed(n + k)
    a = torch.randn(128, k, generator=g)
    b = torch.randn(n, k, generator=g)
    d0 = torch.randn(128, n, generator=g) if flags & 4 else torch.zeros(128, n)
    ref = (a.double() @ b.double().t()) + d0.double()
    scale = float(ref.abs().max())
    ad, bd = a.to(DEV), b.to(DEV)  # keep the device copies alive (the caching allocator would recycle temporaries)
    for passes, tol in ((1, 3e-3), (3, 2e-6), (6, 5e-6), (2, 5e-6)):  # 2 = fp16 x 2 (the fused LFA kernels' arithmetic)
        if passes in (1, 3) and flags & 11:
            continue  # tf32 operands have no no-swizzle MN-major reading (tc.cuh)
        if flags & 8 and k % 32:
            continue
        d = d0.to(DEV) if flags & 4 else torch.full((128, n), float("nan"), device=DEV)
        status = torch.zeros(1, dtype=torch.int32, device=DEV)
        rc = lib.b200_tc_gemm_selftest(c_void_p(ad.data_ptr()), c_void_p(bd.data_ptr()), c_void_p(d.data_ptr()),
                                       n, k, passes, flags, c_void_p(status.data_ptr()),
                                       c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.b200_last_error()
        torch.cuda.synchronize()
        assert int(status) == 0, "tcgen05 completion barrier timed out"
        err = float((d.double().cpu() - ref).abs().max())
        assert err <= tol * scale, f"passes={passes} flags={flags}: max err {err:.3e} vs scale {scale:.3e}"


# ------------------------------------------------------------------------------ loss
@pytest.mark.parametrize("n,c,weighted", [(5000, 6, False), (5000, 7, True), (3, 2, False), (70000, 32, True), (1, 6, False)])
def test_cross_entropy_matches_torch(lib, n, c, weighted):
    """ops.cross_entropy == torch.nn.CrossEntropyLoss(weight, ignore_index=65) (configs/model/criterion/*.yaml) in value
    and gradient, with ignored rows (artefacts are mapped to 65, models/model.py:117-118)."""
    import torch.nn.functional as F

    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(n + c)
    logits = (torch.randn(n, c, generator=g) * 3).to(DEV).requires_grad_(True)
    target = torch.randint(0, c, (n,), generator=g)
    if n > 2:
        target[torch.rand(n, generator=g) < 0.2] = 65
    target = target.to(DEV)
    weight = (torch.rand(c, generator=g) + 0.5).to(DEV) if weighted else None
    ref_in = logits.detach().double().requires_grad_(True)
    ref = F.cross_entropy(ref_in, target, weight=weight.double() if weighted else None, ignore_index=65)
    (ref * 1.7).backward()
    loss = ops.cross_entropy(logits, target, weight, ignore_index=65)
    (loss * 1.7).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert_close(logits.grad, ref_in.grad, atol=1e-7 + 2e-6 * float(ref_in.grad.abs().max()), what="dCE/dlogits")
    assert not logits.grad[target == 65].any()


def test_cross_entropy_all_ignored_is_nan_and_model_uses_it(lib):
    from myria3d_b200 import ops
    from myria3d_b200.model import Model

    logits = torch.randn(10, 6, device=DEV)
    target = torch.full((10,), 65, device=DEV)
    assert torch.isnan(ops.cross_entropy(logits, target, None, 65))  # torch: 0 / 0
    m = Model(neural_net_class_name="B200RandLANet",
              neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
              criterion=torch.nn.CrossEntropyLoss(ignore_index=65), lr=1e-3)
    t = torch.randint(0, 6, (10,), device=DEV)
    want = torch.nn.functional.cross_entropy(logits, t, ignore_index=65)
    assert abs(float(m._loss(logits, t)) - float(want)) < 1e-6
    m.criterion = torch.nn.CrossEntropyLoss(ignore_index=65, label_smoothing=0.1)  # not the fused form: called as given
    assert abs(float(m._loss(logits, t)) - float(m.criterion(logits, t))) == 0.0
