"""End-to-end parity: B200RandLANet (CUDA) vs the CPU oracle of PyGRandLANet.

Mirrors the reference's own tests (tests/myria3d/models/modules/test_randla_nets.py:8-40: shapes for
[12500,12500] / [50,50] / [12500,10000]; tests/myria3d/models/test_model.py:32-53: Model.forward) and adds
the numerical comparison the reference lacks: per-point logits within 1e-3 (fp32, BASELINE north_star),
loss, gradients and BatchNorm running statistics, with the decimation subsets and the dropout mask
injected on both sides.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import randla_oracle as O
from tests.helpers import assert_close, rand_cloud, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOGIT_TOL = 1e-3  # BASELINE.json north_star: per-point logits within 1e-3 fp32


def _pair(num_features=9, num_classes=6, k=16, seed=0, randomize_bn=True):
    from myria3d_b200 import B200RandLANet

    torch.manual_seed(seed)
    ref = O.OracleRandLANet(num_features, num_classes, num_neighbors=k, return_logits=True, knn_method="brute")
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.data.uniform_(0.7, 1.3, generator=g)
                m.bias.data.uniform_(-0.2, 0.2, generator=g)
                m.running_mean.uniform_(-0.2, 0.2, generator=g)
                m.running_var.uniform_(0.6, 1.4, generator=g)
    net = B200RandLANet(num_features, num_classes, num_neighbors=k, return_logits=True)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net.to(DEV)


def _run_train_parity(sizes, k=16, num_classes=6, seed=0, check_grads=True):
    """Train-mode step on both sides.  Ground truth = the oracle evaluated in fp64; the CUDA path must be
    within LOGIT_TOL (1e-3) of it, or -- for batches that are ill-conditioned by construction, e.g. the
    reference's [50, 50] test whose deep levels normalise TWO rows per BatchNorm and amplify fp32 round-off
    by up to 1/sqrt(eps) per layer -- within 10x the error the fp32 oracle (= the reference's own fp32
    arithmetic) itself makes against fp64."""
    ref, net = _pair(num_classes=num_classes, k=k, seed=seed)
    x, pos, batch, ptr = rand_cloud(sizes, seed=seed)
    n = sum(sizes)
    y = torch.randint(0, num_classes, (n,), generator=torch.Generator().manual_seed(seed))
    mask = (torch.rand(n, 32, generator=torch.Generator().manual_seed(seed + 5)) < 0.5).float() * 2.0

    ref64 = O.OracleRandLANet(9, num_classes, num_neighbors=k, return_logits=True, knn_method="brute").double()
    ref64.load_state_dict({k_: (v.double() if v.is_floating_point() else v) for k_, v in ref.state_dict().items()})

    ref.train()
    ref.mlp_classif.injected_masks = [None, mask]
    logits_ref = ref(x, pos, batch, ptr)  # draws its own decimation subsets
    loss_ref = F.cross_entropy(logits_ref, y, ignore_index=65)
    loss_ref.backward()

    ref64.train()
    ref64.mlp_classif.injected_masks = [None, mask.double()]
    logits64 = ref64(x.double(), pos.double(), batch, ptr, decimation_idx=ref.last_decimation_idx)
    loss64 = F.cross_entropy(logits64, y, ignore_index=65)
    loss64.backward()

    net.train()
    net.injected_decimation_idx = ref.last_decimation_idx
    net.mlp_classif.injected_masks = [None, mask.to(DEV)]
    logits = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    loss = F.cross_entropy(logits, y.to(DEV), ignore_index=65)
    loss.backward()

    assert logits.shape == (n, num_classes)
    noise = float((logits_ref.detach().double() - logits64.detach()).abs().max())  # fp32 reference vs fp64
    tol = max(LOGIT_TOL, 10.0 * noise)
    print(f"sizes {sizes}: fp32-oracle noise {noise:.2e}, tolerance {tol:.2e}")
    assert_close(logits, logits64, atol=tol, what=f"train logits {sizes}")
    assert abs(float(loss) - float(loss64)) < max(1e-4, 10 * abs(float(loss_ref) - float(loss64)))
    # (ill-conditioned batches -- 2-row BatchNorms in the reference's [50, 50] case -- are NOT exempt from the gradient
    # check: every parameter gradient must be within max(1e-3, 10 x the fp32 oracle's own error) of the fp64 gradient)
    if check_grads:
        g64 = {k_: p.grad for k_, p in ref64.named_parameters()}
        g32 = {k_: p.grad for k_, p in ref.named_parameters()}
        # yardstick for an ill-conditioned batch: the WORST error the fp32 oracle (= the reference's own arithmetic)
        # makes on any parameter gradient of this batch (round-off amplification is chaotic per parameter)
        e32_worst = max(rel_err(g32[k_], g64[k_]) for k_ in g64 if float(g64[k_].abs().max()) > 1e-6)
        worst = ("", 0.0)
        for name, p in net.named_parameters():
            assert p.grad is not None, f"no grad for {name}"
            r = g64[name]
            e = rel_err(p.grad, r)
            e32 = rel_err(g32[name], r)
            small = float((p.grad.cpu().double() - r).abs().max()) < 2e-6
            if name.endswith(".bias") and (".lins." in name) and "attention.lins" not in name.replace("post_attention", ""):
                # Linear bias in front of a train-mode BatchNorm: gradient is mathematically zero
                wref = g64[name.replace("bias", "weight")]
                small = small or float(p.grad.abs().max()) <= 1e-3 * float(wref.abs().max()) + 1e-6
            if e > worst[1] and not small:
                worst = (name, e)
            # rel 1e-3 on every parameter gradient, relaxed like the logits when ill-conditioned
            assert e < max(1e-3, 10 * e32, 10 * e32_worst if noise > 1e-4 else 0.0) or small, \
                f"grad {name}: rel err {e:.3e} (fp32 oracle: {e32:.3e}, its worst parameter: {e32_worst:.3e})"
        print("worst grad rel err", worst)
    ref_bufs = dict(ref64.named_buffers())
    for name, b in net.named_buffers():
        assert_close(b, ref_bufs[name], atol=max(1e-5, 10 * noise * 0.01), rtol=1e-4, what=f"buffer {name}")
    return ref, net


@pytest.mark.parametrize("sizes", [[50, 50], [1250, 1000], [700, 3, 40, 2000]])
def test_train_step_parity(lib, sizes):
    _run_train_parity(sizes)


def test_train_step_parity_k32(lib):
    _run_train_parity([900, 40], k=32)


def test_train_step_parity_baseline_tile(lib):
    """Train-mode NUMERICAL parity at the BASELINE tile size (the reference's own test only checks shapes at
    [12500, 12500], tests/myria3d/models/modules/test_randla_nets.py:8-40): one 12 800-point tile + a 3 000-point one.
    Exercises what the small cases cannot: the tcgen05 GEMMs with their BatchNorm-statistics epilogue (n >= 1024 rows
    at >= 64 channels), the warp-cooperative grid kNN on levels 0-1, the tensor-core LFA forward / backward with many
    tiles per slot and the periodic dW flush.  Ground truth: the fp64 oracle."""
    _run_train_parity([12800, 3000], seed=3)


@pytest.mark.parametrize("num_nodes", [[12500, 12500], [50, 50], [12500, 10000]])
def test_fake_run_b200_randlanet(lib, num_nodes):
    """The reference's shape test (test_randla_nets.py:8-40), default train mode, log-prob output."""
    from myria3d_b200 import B200RandLANet

    x, pos, batch, ptr = rand_cloud(num_nodes, seed=11)
    model = B200RandLANet(9, 6, decimation=4, num_neighbors=16).to(DEV)
    out = model(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    assert out.shape == torch.Size([sum(num_nodes), 6])
    assert torch.isfinite(out).all()
    assert_close(out.exp().sum(1), torch.ones(sum(num_nodes)), atol=1e-4, what="log_softmax rows")


def test_eval_parity_synthetic_tile(lib):
    """Eval-mode forward (running statistics) on Lidar-HD-like synthetic tiles, config-A sized (2 x 4096)."""
    ref, net = _pair(seed=3)
    x, pos, y, batch, ptr = O.synthetic_batch([4096, 4096], seed=12345)
    ref.eval(), net.eval()
    with torch.no_grad():
        logits_ref = ref(x, pos, batch, ptr)
        net.injected_decimation_idx = ref.last_decimation_idx
        logits = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    assert_close(logits, logits_ref, atol=LOGIT_TOL, what="eval logits")
    assert (logits.argmax(1).cpu() == logits_ref.argmax(1)).float().mean() > 0.999


def test_single_point_cloud_eval(lib):
    """tests/myria3d/test_train_and_predict.py:130-143: a one-point cloud must go through in eval mode."""
    ref, net = _pair(seed=4)
    x, pos, batch, ptr = rand_cloud([1], seed=5)
    ref.eval(), net.eval()
    with torch.no_grad():
        a = ref(x, pos, batch, ptr)
        net.injected_decimation_idx = ref.last_decimation_idx
        b = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    assert_close(b, a, atol=LOGIT_TOL, what="single point logits")


def test_decimation_matches_reference_rng_stream(lib):
    """Same torch.randperm calls in the same order as pyg_randla_net.py:219-229 on the same device."""
    from myria3d_b200.randla_net import decimation_indices

    ptr = [0, 100, 103, 1103]
    torch.manual_seed(5)
    idx, new_ptr = decimation_indices(ptr, 4, torch.device(DEV))
    torch.manual_seed(5)
    expect = torch.cat([ptr[i] + torch.randperm(ptr[i + 1] - ptr[i], device=DEV)[: max(1, (ptr[i + 1] - ptr[i]) // 4)]
                        for i in range(3)])
    assert torch.equal(idx, expect) and new_ptr == [0, 25, 26, 276]


def test_checkpoint_state_dict_and_model_wrapper(lib):
    """Model wrapper surface (models/model.py:67-120): forward -> (targets, logits); training_step dict."""
    from myria3d_b200 import Batch, Data, Model

    model = Model(neural_net_class_name="B200RandLANet",
                  neural_net_hparams=dict(num_features=9, num_classes=7, num_neighbors=16, decimation=4, return_logits=True),
                  criterion=torch.nn.CrossEntropyLoss(ignore_index=65), interpolation_k=10, num_workers=1).to(DEV)
    datas = []
    for i, n in enumerate([600, 300]):
        x, pos, y = O.synthetic_tile(n, 100 + i, num_classes=7)
        datas.append(Data(x=x, pos=pos, y=y))
    batch = Batch.from_data_list(datas).to(DEV)
    model.train()
    out = model.training_step(batch, 0)
    assert set(out) == {"loss", "logits", "targets"} and out["logits"].shape == (900, 7)
    out["loss"].backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())

    # eval with `copies`: interpolation of logits to the full cloud (model.py:86-98) vs the oracle on CPU
    model.eval()
    full = [torch.rand(2000, 3), torch.rand(900, 3)]
    sub = [d.pos for d in datas]
    batch = Batch.from_data_list(datas)
    batch.copies = {"pos_copy": torch.cat(full), "pos_sampled_copy": torch.cat(sub)}
    batch.idx_in_original_cloud = [torch.arange(2000).numpy(), torch.arange(900).numpy()]
    with torch.no_grad():
        targets, logits_full = model(batch.to(DEV))
    assert targets is None and logits_full.shape == (2900, 7)
    # the GPU interpolation (b200_knn k = 10 + b200_knn_interp) == PyG knn_interpolate(k = interpolation_k) of
    # model.py:90-98 applied to the same sub-sampled logits, cloud by cloud
    model.model.injected_decimation_idx = [t.clone() for t in model.model.last_decimation_idx]  # eval decimates too
    with torch.no_grad():
        logits_sub = model.model(batch.x.to(DEV), batch.pos.to(DEV), None, batch.ptr.to(DEV)).cpu()
    model.model.injected_decimation_idx = None
    expect = O.knn_interpolate(logits_sub, torch.cat(sub), torch.cat(full), [0, 600, 900], [0, 2000, 2900], 10, method="brute")
    assert_close(logits_full, expect, atol=2e-5 * float(expect.abs().max()), what="eval logits interpolated to the full cloud")
    pred = model.predict_step(batch.to(DEV))
    assert pred["logits"].device.type == "cpu" and pred["logits"].shape == (2900, 7)


def test_against_committed_golden_vectors(lib):
    """CUDA path vs tests/golden/randla_small.pt (oracle outputs committed by oracle/gen_golden.py)."""
    import os

    from myria3d_b200 import B200RandLANet
    from oracle.gen_golden import build_net

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "randla_small.pt"))
    net = B200RandLANet(9, 6, num_neighbors=16, return_logits=True)
    net.load_state_dict(build_net(g["seed"]).state_dict(), strict=True)
    net.to(DEV)
    args = [g[k].to(DEV) for k in ("x", "pos", "batch", "ptr")]
    net.injected_decimation_idx = g["decimation_idx"]

    from myria3d_b200 import ops

    nbr, _ = ops.knn(args[1], args[3], args[1], args[3], 16, max(g["sizes"]), kt=16)
    assert torch.equal(nbr.cpu(), g["knn_level0"]), "level-0 kNN table differs from the golden one"

    net.eval()
    with torch.no_grad():
        logits = net(*args)
    assert_close(logits, g["logits_eval"], atol=LOGIT_TOL, what="eval logits vs golden")

    net.load_state_dict(build_net(g["seed"]).state_dict(), strict=True)
    net.train()
    net.mlp_classif.injected_masks = [None, g["dropout_mask"].to(DEV)]
    logits = net(*args)
    loss = F.cross_entropy(logits, g["y"].to(DEV), ignore_index=65)
    loss.backward()
    assert_close(logits, g["logits_train"], atol=LOGIT_TOL, what="train logits vs golden")
    assert abs(float(loss.detach()) - g["loss"]) < 1e-4
    params = dict(net.named_parameters())
    for k, v in g["grads"].items():
        assert rel_err(params[k].grad, v) < 1e-3, f"grad {k}"
    bufs = dict(net.named_buffers())
    for k, v in g["buffers_after_step"].items():
        assert_close(bufs[k], v, atol=1e-5, rtol=1e-4, what=k)


def test_param_grads_accumulate_into_existing_buffers(lib):
    """Second backward on the same batch: every parameter already owns a dense .grad, so the kernels accumulate straight
    into it (ops._direct_grad: Linear, BatchNorm affine, encoder fold, attention weights) -- the result must be exactly
    twice... up to fp32 rounding of g + g == 2g, i.e. exactly 2x the first gradient wherever the first pass is
    deterministic, and within 1e-5 relative where atomics reorder sums."""
    ref, net = _pair(seed=21)
    x, pos, batch, ptr = rand_cloud([900, 400], seed=21)
    y = torch.randint(0, 6, (1300,), generator=torch.Generator().manual_seed(1)).to(DEV)
    net.train()
    net.mlp_classif.injected_masks = [None, (torch.rand(1300, 32, device=DEV) < 0.5).float() * 2.0]  # same dropout twice
    args = (x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    out = net(*args)
    idx = [t.clone() for t in net.last_decimation_idx]
    F.cross_entropy(out, y).backward()
    first = {n_: p.grad.clone() for n_, p in net.named_parameters()}
    assert all(p.grad is not None for p in net.parameters())
    net.injected_decimation_idx = idx
    out = net(*args)  # batch statistics are identical; only the running buffers moved
    F.cross_entropy(out, y).backward()
    top = max(float(g.abs().max()) for g in first.values())
    for n_, p in net.named_parameters():
        scale = float(first[n_].abs().max())
        # (biases in front of a BatchNorm have a mathematically zero gradient: fp32 noise, different on every run)
        assert_close(p.grad, 2 * first[n_], atol=2e-5 * scale + 1e-6 * top, what=f"accumulated grad of {n_}")


def test_against_trained_checkpoint_golden(lib):
    """CUDA path under the reference's shipped, TRAINED weights (tests/golden/randla_trained_ckpt.pt, generated by
    oracle/gen_golden_ckpt.py from trained_model_assets/proto151_V2.0_epoch_100_Myria3DV3.1.0.ckpt): strict state-dict
    load, eval-mode logits within 1e-3 of the fp64 oracle (logit magnitudes up to 22 here; the fp32 oracle itself is
    4e-5 away) and identical predicted classes."""
    import os

    from myria3d_b200 import B200RandLANet
    from myria3d_b200.synthetic import synthetic_batch

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "randla_trained_ckpt.pt"))
    net = B200RandLANet(g["num_features"], g["num_classes"], num_neighbors=g["k"], return_logits=True)
    net.load_state_dict(g["state_dict"], strict=True)
    net.to(DEV).eval()
    x, pos, _, batch, ptr = synthetic_batch(g["sizes"], seed=g["seed"], num_features=9, num_classes=7)
    net.injected_decimation_idx = g["decimation_idx"]
    with torch.no_grad():
        logits = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    tol = max(LOGIT_TOL, 10 * g["fp32_vs_fp64_max_err"])
    assert_close(logits, g["logits_fp64"], atol=tol, what="eval logits vs trained-checkpoint golden (fp64 oracle)")
    agree = (logits.argmax(1).cpu() == g["logits_fp64"].argmax(1)).float().mean()
    assert float(agree) >= 0.999, float(agree)


def test_config_a_block1_net_parity(lib):
    """BASELINE.json configs[0] ('1 encoder layer', 4 096 points x batch 2, 6 classes): the product's block-1 network
    (fc0 + block1 + head) against the oracle's, train mode -- logits within 1e-3 of fp64, every parameter gradient
    within max(1e-3, 10 x fp32-oracle error), BatchNorm buffers -- and eval mode."""
    from myria3d_b200 import B200Block1Net, get_neural_net_class

    assert get_neural_net_class("B200Block1Net") is B200Block1Net
    torch.manual_seed(0)
    ref = O.OracleBlock1Net(9, 6, knn_method="brute")
    ref64 = O.OracleBlock1Net(9, 6, knn_method="brute").double()
    ref64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in ref.state_dict().items()})
    net = B200Block1Net(9, 6)
    net.load_state_dict(ref.state_dict(), strict=True)
    net.to(DEV)
    x, pos, y, batch, ptr = O.synthetic_batch([4096, 4096], seed=12345)
    mask = (torch.rand(8192, 32, generator=torch.Generator().manual_seed(5)) < 0.5).float() * 2.0
    for m_, mk in ((ref, mask), (ref64, mask.double()), (net, mask.to(DEV))):
        m_.train()
        m_.mlp_classif.injected_masks = [None, mk]
    lr_ = ref(x, pos, batch, ptr)
    F.cross_entropy(lr_, y).backward()
    l64 = ref64(x.double(), pos.double(), batch, ptr)
    F.cross_entropy(l64, y).backward()
    lg = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
    F.cross_entropy(lg, y.to(DEV)).backward()
    noise = float((lr_.detach().double() - l64.detach()).abs().max())
    assert_close(lg, l64, atol=max(LOGIT_TOL, 10 * noise), what="config A train logits")
    g64, g32 = dict(ref64.named_parameters()), dict(ref.named_parameters())
    for name, p in net.named_parameters():
        r = g64[name].grad
        e, e32 = rel_err(p.grad, r), rel_err(g32[name].grad, r)
        tiny = float((p.grad.cpu().double() - r).abs().max()) < 2e-6
        assert e < max(1e-3, 10 * e32) or tiny, f"grad {name}: rel err {e:.3e} (fp32 oracle {e32:.3e})"
    bufs = dict(ref64.named_buffers())
    for name, b in net.named_buffers():
        assert_close(b, bufs[name], atol=1e-5, rtol=1e-4, what=name)
    ref.eval(), net.eval()
    with torch.no_grad():
        assert_close(net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV)), ref(x, pos, batch, ptr), atol=LOGIT_TOL,
                     what="config A eval logits")


def test_eval_logits_against_reference_model_code_vectors(lib):
    """CUDA eval forward against logits computed by the reference's OWN model file on stand-in PyG primitives
    (tests/golden/ref_model_standin.pt; oracle/gen_golden_ref_model.py).  The decimation subsets are the reference's
    randperm draws, recovered by re-running the (bit-identical, see test_oracle_equals_reference_model_code) oracle
    on the same seed."""
    import os
    from myria3d_b200 import B200RandLANet

    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_standin.pt"))
    for name, c in gold["cases"].items():
        sizes, k, seed = c["sizes"], c["k"], c["seed"]
        torch.manual_seed(seed)
        ref = O.OracleRandLANet(9, 6, num_neighbors=k, return_logits=True)
        g = torch.Generator().manual_seed(seed + 1)
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.data.uniform_(0.7, 1.3, generator=g)
                m.bias.data.uniform_(-0.2, 0.2, generator=g)
                m.running_mean.uniform_(-0.2, 0.2, generator=g)
                m.running_var.uniform_(0.6, 1.4, generator=g)
        x, pos, y, batch, ptr = O.synthetic_batch(sizes, seed=seed)
        ref.eval()
        torch.manual_seed(seed + 2)
        with torch.no_grad():
            lr = ref(x, pos, batch, ptr)
        assert torch.equal(lr, gold[name]["eval_logits"])  # the oracle run IS the reference run
        net = B200RandLANet(9, 6, num_neighbors=k, return_logits=True)
        net.load_state_dict(ref.state_dict(), strict=True)
        net = net.to(DEV).eval()
        net.injected_decimation_idx = ref.last_decimation_idx
        with torch.no_grad():
            lg = net(x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV))
        assert_close(lg, gold[name]["eval_logits"], atol=LOGIT_TOL, what=f"eval logits vs reference model code ({name})")
