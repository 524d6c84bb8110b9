import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from myria3d_b200 import B200RandLANet, _lib, ops
from oracle import randla_oracle as O
for seed in (7, 8, 9, 10, 11, 12, 13):
    torch.manual_seed(12345)
    ref = O.OracleRandLANet(9, 6, num_neighbors=16, return_logits=True, knn_method="brute")
    net = B200RandLANet(9, 6, num_neighbors=16, return_logits=True)
    net.load_state_dict(ref.state_dict(), strict=True)
    net = net.to("cuda:0")
    x, pos, y, batch, ptr = O.synthetic_batch([1500, 600], seed=seed)
    mask = (torch.rand(2100, 32) < 0.5).float() * 2.0
    ref.train(), net.train()
    ref.mlp_classif.injected_masks = [None, mask]
    logits_ref = ref(x, pos, batch, ptr)
    F.cross_entropy(logits_ref, y, ignore_index=65).backward()
    net.injected_decimation_idx = ref.last_decimation_idx
    net.mlp_classif.injected_masks = [None, mask.cuda()]
    logits = net(x.cuda(), pos.cuda(), batch.cuda(), ptr.cuda())
    ops.cross_entropy(logits, y.cuda(), None, ignore_index=65).backward()
    gref = ref.fc0.weight.grad
    errs = sorted(float((p.grad.cpu() - q.grad).norm() / (q.grad.norm() + 1e-30)) for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()) if "lins.0.bias" not in n and "lins.1.bias" not in n and n != "fc0.bias")
    print(f"seed {seed}: logits {float((logits.detach().cpu() - logits_ref.detach()).abs().max()):.2e} fc0.weight {float((net.fc0.weight.grad.cpu() - gref).norm() / gref.norm()):.2e} median {errs[len(errs)//2]:.2e} max {errs[-1]:.2e}")
