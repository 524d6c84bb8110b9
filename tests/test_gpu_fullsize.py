"""BASELINE.json configs[3] / configs[4] at their full sizes, checked through size-independent properties (the CPU
oracle needs minutes at these sizes): exactness of the grid k-NN against the brute-force kernel (itself pinned bit-exact
against the oracle at small sizes), determinism, independence of the tiles of a batch, finite gradients, and the
predict-time chain forward -> k=10 interpolation -> sliding-window stitch against the CPU scatter."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import randla_oracle as O
from oracle import stitch_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _net(k, classes=6, seed=0):
    from myria3d_b200 import B200RandLANet

    torch.manual_seed(seed)
    return B200RandLANet(9, classes, num_neighbors=k, return_logits=True).to(DEV)


def test_config_e_k32_65536_points(lib):
    """configs[4]: K=32, 4 tiles x 65 536 points."""
    from myria3d_b200 import ops
    from myria3d_b200.randla_net import _Level

    sizes = [65536] * 4
    x, pos, y, batch, ptr = O.synthetic_batch(sizes, seed=777)
    xd, posd, bd, ptrd, yd = x.to(DEV), pos.to(DEV), batch.to(DEV), ptr.to(DEV), y.to(DEV)

    # k-NN: grid search == brute force, bit for bit; rows ascending in distance, self first
    lvl = _Level(ptr.tolist(), torch.device(DEV))
    nbr_g, d_g = ops.knn(posd, lvl.ptr, posd, lvl.ptr, 32, lvl.max_n, kt=32, algo="grid")
    nbr_b, d_b = ops.knn(posd, lvl.ptr, posd, lvl.ptr, 32, lvl.max_n, kt=32, algo="brute")
    assert torch.equal(nbr_g, nbr_b) and torch.equal(d_g, d_b)
    assert torch.equal(nbr_g[:, 0].long(), torch.arange(sum(sizes), device=DEV))
    assert (d_g[:, 1:] >= d_g[:, :-1]).all() and (d_g[:, 0] == 0).all()
    cloud_of = bd[nbr_g.long()]
    assert (cloud_of == bd[:, None]).all()  # neighbours never cross tiles

    # eval forward: deterministic, and tile 0 of the batch == tile 0 alone (same subsets) to fp32 round-off
    net = _net(32).eval()
    with torch.no_grad():
        a = net(xd, posd, bd, ptrd)
        idx = [t.clone() for t in net.last_decimation_idx]
        net.injected_decimation_idx = idx
        b = net(xd, posd, bd, ptrd)
        assert torch.equal(a, b)  # no atomics on the forward path
        n0 = sizes[0]
        idx0, n_l = [], n0
        for t in idx:
            n_l = max(1, n_l // 4)
            idx0.append(t[:n_l])
        net.injected_decimation_idx = idx0
        alone = net(xd[:n0], posd[:n0], bd[:n0], ptrd[:2])
    assert torch.isfinite(a).all()
    assert float((alone - a[:n0]).abs().max()) <= 1e-4 * max(1.0, float(a.abs().max()))

    # train step: finite loss and gradients for every parameter
    net.train()
    net.injected_decimation_idx = None
    loss = F.cross_entropy(net(xd, posd, bd, ptrd), yd)
    loss.backward()
    assert torch.isfinite(loss)
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name


def test_config_d_predict_chain_40960_point_tiles(lib):
    """configs[3]: 40 960-point receptive fields subsampled from denser windows, eval forward, k=10 inverse-distance
    interpolation back to every point of the window (models/model.py:86-98), overlapping windows stitched into the full
    cloud (models/interpolation.py:98-121)."""
    from myria3d_b200 import Batch, Data, Model
    from myria3d_b200.interpolation import Interpolator

    classes = {1: "unclassified", 2: "ground", 6: "building", 9: "water", 17: "bridge", 64: "lasting_above"}
    full_per_window, sub, n_windows = 60000, 40960, 3
    nb_points = 150000  # the full cloud; consecutive windows share 15 000 points
    model = Model(neural_net_class_name="B200RandLANet",
                  neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
                  criterion=torch.nn.CrossEntropyLoss(ignore_index=65), interpolation_k=10, num_workers=1).to(DEV)
    model.eval()
    g = torch.Generator().manual_seed(99)
    datas = []
    for w in range(n_windows):
        x, pos, y = O.synthetic_tile(full_per_window, seed=500 + w)
        keep = torch.randperm(full_per_window, generator=g)[:sub]
        d = Data(x=x[keep], pos=pos[keep], y=y[keep])
        d.copies = {"pos_copy": pos.clone(), "pos_sampled_copy": pos[keep].clone(), "transformed_y_copy": y.clone()}
        d.idx_in_original_cloud = np.arange(w * 45000, w * 45000 + full_per_window, dtype=np.int64)
        datas.append((d, keep))
    batch = Batch.from_data_list([d for d, _ in datas]).to(DEV)  # Lightning moves the batch to the module's device
    with torch.no_grad():
        out = model.predict_step(batch)
    logits = out["logits"]
    assert logits.shape == (n_windows * full_per_window, 6) and torch.isfinite(logits).all()

    # a point that was kept by the subsampling is its own nearest neighbour at distance 0: weight 1e16 dominates the
    # other nine, so the interpolated logits equal the network's logits there
    model.model.injected_decimation_idx = [t.clone() for t in model.model.last_decimation_idx]  # same random subsets
    with torch.no_grad():
        sub_logits = model.model(batch.x, batch.pos, batch.batch, batch.ptr).cpu()
    for w, (_, keep) in enumerate(datas):
        got = logits[w * full_per_window:(w + 1) * full_per_window][keep]
        want = sub_logits[w * sub:(w + 1) * sub]
        assert float((got - want).abs().max()) <= 1e-3 * max(1.0, float(want.abs().max()))

    # stitch: GPU Interpolator == CPU restatement (bit-exact sums, preds; 1e-5 entropy)
    per_window = [logits[w * full_per_window:(w + 1) * full_per_window] for w in range(n_windows)]
    idx_list = [d.idx_in_original_cloud for d, _ in datas]
    want, want_idx = SO.reduce_predictions(per_window, idx_list, nb_points, classes)
    itp = Interpolator(interpolation_k=10, classification_dict=classes)
    for l, i in zip(per_window, idx_list):
        itp.store_predictions(l.to(DEV), [i])
    got, got_idx = itp.reduce_predictions(nb_points)
    assert np.array_equal(got_idx, want_idx)
    assert np.array_equal(got["PredictedClassification"], want["preds"])
    np.testing.assert_allclose(got["entropy"], want["entropy"].numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got["ground"], want["probas"][:, 1].numpy(), rtol=2e-6, atol=1e-7)
    # overlap really happened: the shared points got two predictions
    counts = np.bincount(want_idx, minlength=nb_points)
    assert counts.max() == 2 and (counts == 2).sum() == 2 * 15000
