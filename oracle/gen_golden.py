"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

    python -m oracle.gen_golden

TEST INFRASTRUCTURE.  The reference has no numerical fixtures for this path and its
third-party stack cannot be imported here, so these vectors pin the ORACLE (regression guard + a
ground truth the GPU tests can use without recomputing); they are not outputs of the reference itself.
Weights are re-derived from the seed (a checksum is stored), inputs/outputs are stored in full.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from oracle import randla_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 12345  # configs/config.yaml:3, tests/conftest.py:43-45 of the reference


def build_net(seed: int = SEED, num_classes: int = 6, k: int = 16):
    torch.manual_seed(seed)
    net = O.OracleRandLANet(9, num_classes, num_neighbors=k, return_logits=True, knn_method="brute")
    g = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.uniform_(0.7, 1.3, generator=g)
            m.bias.data.uniform_(-0.2, 0.2, generator=g)
            m.running_mean.uniform_(-0.2, 0.2, generator=g)
            m.running_var.uniform_(0.6, 1.4, generator=g)
    return net


def weight_checksum(net) -> float:
    return float(sum(p.double().abs().sum() for p in net.state_dict().values() if p.is_floating_point()))


def main():
    os.makedirs(OUT, exist_ok=True)
    sizes = [800, 500]  # summit BatchNorm sees 5 rows; [300, 77] (2 rows) amplifies fp32 noise 100x
    x, pos, y, batch, ptr = O.synthetic_batch(sizes, seed=SEED)
    n = sum(sizes)
    mask = (torch.rand(n, 32, generator=torch.Generator().manual_seed(SEED + 5)) < 0.5).float() * 2.0

    net = build_net()
    checksum = weight_checksum(net)
    net.eval()
    with torch.no_grad():
        logits_eval = net(x, pos, batch, ptr)
    dec_idx = [t.clone() for t in net.last_decimation_idx]
    knn0 = O.knn_bruteforce(pos, ptr.tolist(), pos, ptr.tolist(), 16)[0]

    net = build_net()
    net.train()
    net.mlp_classif.injected_masks = [None, mask]
    logits_train = net(x, pos, batch, ptr, decimation_idx=dec_idx)
    loss = F.cross_entropy(logits_train, y, ignore_index=65)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()
             if k in ("fc0.weight", "fc_classif.weight", "block1.lfa1.mlp_attention.lins.0.weight",
                      "block4.lfa2.mlp_encoder.lins.0.weight", "fp2.nn.lins.0.weight",
                      "block2.mlp2.norms.0.module.weight")}
    bufs = {k: v.clone() for k, v in net.named_buffers()
            if k in ("block1.lfa1.mlp_encoder.norms.0.module.running_var", "mlp_summit.norms.0.module.running_mean")}
    torch.save({
        "sizes": sizes, "seed": SEED, "x": x, "pos": pos, "y": y, "batch": batch, "ptr": ptr,
        "dropout_mask": mask, "decimation_idx": dec_idx, "knn_level0": knn0.int(),
        "weight_checksum": checksum, "logits_eval": logits_eval, "logits_train": logits_train.detach(),
        "loss": float(loss), "grads": grads, "buffers_after_step": bufs,
        "torch_version": str(torch.__version__),
    }, os.path.join(OUT, "randla_small.pt"))
    print("wrote", os.path.join(OUT, "randla_small.pt"), "checksum", checksum, "loss", float(loss))


if __name__ == "__main__":
    main()
