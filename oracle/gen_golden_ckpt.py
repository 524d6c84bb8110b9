"""Golden vectors under the reference's SHIPPED, TRAINED weights (tests/golden/randla_trained_ckpt.pt).

    python -m oracle.gen_golden_ckpt        # needs /root/reference (this container only)

TEST INFRASTRUCTURE.  SURVEY.md 8c: the only numerical artefact the reference ships for this path is its trained
checkpoint (``trained_model_assets/proto151_V2.0_epoch_100_Myria3DV3.1.0.ckpt``: 257 state entries, F = 9, C = 7).
The fixture stores those weights (they cannot travel otherwise: /root/reference does not exist on the GPU box), the
decimation subsets of one eval-mode forward of the ORACLE on two synthetic Lidar-HD-like tiles (inputs are re-derived
from the seed), and the oracle's logits in fp32 and in fp64.  It pins (a) the oracle against itself (regression) and
(b) the CUDA path against the oracle under realistic weights, BatchNorm running statistics and logit magnitudes
(|logit| up to 24; random-init tests stay near 1).  Since round 2 the generator also runs the reference's OWN model file (on the stand-in PyG primitives of
oracle/pyg_standin.py) with these weights and asserts bit-identical logits: `logits_reference_model_code`.
"""
from __future__ import annotations

import os

import torch

from myria3d_b200.ckpt import load_lightning_checkpoint, net_state_dict
from oracle import randla_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = "/root/reference/trained_model_assets/proto151_V2.0_epoch_100_Myria3DV3.1.0.ckpt"
OUT = os.path.join(ROOT, "tests", "golden", "randla_trained_ckpt.pt")
SIZES, SEED = [4096, 3000], 2024


def main():
    sd = net_state_dict(load_lightning_checkpoint(CKPT))
    net = O.OracleRandLANet(9, 7, num_neighbors=16, return_logits=True, knn_method="brute")
    net.load_state_dict(sd, strict=True)
    net.eval()
    x, pos, y, batch, ptr = O.synthetic_batch(SIZES, seed=SEED, num_features=9, num_classes=7)
    torch.manual_seed(SEED)
    with torch.no_grad():
        logits32 = net(x, pos, batch, ptr)
        idx = [t.clone() for t in net.last_decimation_idx]
        net64 = O.OracleRandLANet(9, 7, num_neighbors=16, return_logits=True, knn_method="brute").double()
        net64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
        net64.eval()
        logits64 = net64(x.double(), pos.double(), batch, ptr, decimation_idx=idx)
    # the same forward through the reference's OWN model file on stand-in PyG primitives (oracle/pyg_standin.py): must
    # be bit-identical to the oracle's -- recorded in the fixture so that the tests holding the CUDA path to `logits_fp32`
    # hold it to the reference's model code under its shipped weights
    from oracle import gen_golden_ref_model as G, pyg_standin

    pyg_standin.KNN_METHOD = "brute"
    ref_net = G.load_reference_model_module().PyGRandLANet(9, 7, decimation=4, num_neighbors=16, return_logits=True)
    ref_net.load_state_dict(sd, strict=True)
    ref_net.eval()
    torch.manual_seed(SEED)
    with torch.no_grad():
        logits_ref_code = ref_net(x, pos, batch, ptr)
    assert torch.equal(logits_ref_code, logits32), float((logits_ref_code - logits32).abs().max())
    torch.save({
        "logits_reference_model_code": logits_ref_code, "reference_model_code_equals_oracle": True,
        "source": os.path.basename(CKPT), "sizes": SIZES, "seed": SEED, "num_features": 9, "num_classes": 7, "k": 16,
        "state_dict": {k: v.clone() for k, v in sd.items()},
        "decimation_idx": idx, "logits_fp32": logits32, "logits_fp64": logits64,
        "fp32_vs_fp64_max_err": float((logits32.double() - logits64).abs().max()),
        "torch_version": str(torch.__version__),
    }, OUT)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1e6:.2f} MB; |logit| max {float(logits32.abs().max()):.2f}; "
          f"fp32 vs fp64 max err {float((logits32.double() - logits64).abs().max()):.2e}")


if __name__ == "__main__":
    main()
