"""Sliding-window stitch: CUDA kernels behind ``myria3d_b200.interpolation.Interpolator`` against the CPU
oracle restatement of ``myria3d/models/interpolation.py``."""
import numpy as np
import pytest
import torch

from oracle import stitch_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda"

CLASSES = {1: "unclassified", 2: "ground", 6: "building", 9: "water", 17: "bridge", 64: "lasting_above", 65: "artefact"}


def windows(nb_points: int, n_windows: int, size: int, seed: int):
    """Overlapping prediction windows: each a random subset of the cloud, so points get 0..n_windows predictions."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    idx_list, logits_list = [], []
    for w in range(n_windows):
        n = min(size, nb_points)
        idx_list.append(rng.choice(nb_points, size=n, replace=False).astype(np.int64))
        logits_list.append(torch.randn(n, len(CLASSES), generator=g) * 3)
    return logits_list, idx_list


@pytest.mark.parametrize("nb_points,n_windows,size", [(1000, 5, 600), (50000, 9, 20000), (7, 4, 7), (300, 1, 100)])
def test_scatter_sum_is_bit_identical_to_the_cpu_scatter(nb_points, n_windows, size):
    from myria3d_b200 import ops

    logits_list, idx_list = windows(nb_points, n_windows, size, seed=nb_points)
    logits = torch.cat(logits_list)
    idx = torch.from_numpy(np.concatenate(idx_list))
    want = SO.scatter_sum_rows(logits, idx, nb_points)
    got = ops.stitch_scatter_sum(logits.to(DEV), idx.to(DEV), nb_points).cpu()
    assert torch.equal(got, want)  # up to 9 contributions per point, summed in input order
    if nb_points <= 1000:
        assert torch.equal(want, SO.scatter_sum_rows_loop(logits, idx.tolist(), nb_points))


def test_scatter_sum_against_reference_run_vectors():
    """CUDA scatter-sum + gather against the output of the REFERENCE'S OWN Interpolator.reduce_predicted_logits
    (tests/golden/ref_sample_prep.npz, oracle/gen_golden_ref.py): bit-exact."""
    import os
    from myria3d_b200 import ops

    ref = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sample_prep.npz"))
    logits, idx = torch.from_numpy(ref["stitch_logits"]).to(DEV), torch.from_numpy(ref["stitch_idx"]).to(DEV)
    summed = ops.stitch_scatter_sum(logits, idx, int(ref["stitch_nb_points"]))
    assert np.array_equal(summed[idx].cpu().numpy(), ref["stitch_reduced"])


def test_scatter_sum_odd_class_count_and_errors():
    from myria3d_b200 import ops

    g = torch.Generator().manual_seed(0)
    logits = torch.randn(500, 5, generator=g)  # c % 4 != 0 -> scalar path
    idx = torch.randint(0, 40, (500,), generator=g)
    want = SO.scatter_sum_rows(logits, idx, 64)
    got = ops.stitch_scatter_sum(logits.to(DEV), idx.to(DEV), 64).cpu()
    assert torch.equal(got, want)
    with pytest.raises(IndexError):
        ops.stitch_scatter_sum(logits.to(DEV), idx.to(DEV), 10)
    empty = ops.stitch_scatter_sum(torch.zeros(0, 7, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 12)
    assert empty.shape == (12, 7) and not empty.any()


@pytest.mark.parametrize("nb_points,n_windows,size", [(2000, 4, 1200), (40000, 3, 30000)])
def test_interpolator_matches_the_reference_restatement(nb_points, n_windows, size):
    from myria3d_b200.interpolation import Interpolator

    logits_list, idx_list = windows(nb_points, n_windows, size, seed=7 + nb_points)
    want, want_idx = SO.reduce_predictions(logits_list, idx_list, nb_points, CLASSES)

    itp = Interpolator(interpolation_k=10, classification_dict=CLASSES, probas_to_save="all")
    for l, i in zip(logits_list, idx_list):
        itp.store_predictions(l.to(DEV), [i])  # predict_step hands one index array per sample
    got_logits, got_idx = itp.reduce_predicted_logits(nb_points)
    assert np.array_equal(got_idx, want_idx)
    assert torch.equal(got_logits.cpu(), want["logits"])  # bit-exact sums and gather

    itp = Interpolator(interpolation_k=10, classification_dict=CLASSES, probas_to_save=["building", "ground"])
    for l, i in zip(logits_list, idx_list):
        itp.store_predictions(l.to(DEV), [i])
    out, got_idx = itp.reduce_predictions(nb_points)
    assert set(out) == {"building", "ground", "PredictedClassification", "entropy"}
    names = list(CLASSES.values())
    for name in ("building", "ground"):
        np.testing.assert_allclose(out[name], want["probas"][:, names.index(name)].numpy(), rtol=2e-6, atol=1e-7)
    assert np.array_equal(out["PredictedClassification"], want["preds"])  # LAS codes, argmax on bit-identical logits
    np.testing.assert_allclose(out["entropy"], want["entropy"].numpy(), rtol=1e-5, atol=2e-6)


def test_finalize_handles_extreme_logits():
    """Saturated softmax (p = 1 / p = 0): entropy clamps like torch's probs_to_logits, no NaN."""
    from myria3d_b200 import ops

    reduced = torch.tensor([[100.0, -100.0, 0.0], [0.0, 0.0, 0.0], [-5.0, 80.0, 80.0]])
    idx = torch.tensor([2, 0, 1, 1])
    _, probas, preds, ent = ops.stitch_finalize(reduced.to(DEV), idx.to(DEV))
    ref_l = reduced[idx]
    ref_p = torch.softmax(ref_l, dim=1)
    assert torch.equal(preds.cpu(), torch.argmax(ref_l, dim=1))  # ties -> first maximum
    torch.testing.assert_close(probas.cpu(), ref_p, rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(ent.cpu(), torch.distributions.Categorical(probs=ref_p).entropy(), rtol=1e-5, atol=2e-6)
    assert torch.isfinite(ent).all()
