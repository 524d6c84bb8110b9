import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from myria3d_b200 import B200RandLANet, _lib, ops
from oracle import randla_oracle as O
torch.manual_seed(12345)
ref = O.OracleRandLANet(9, 6, num_neighbors=16, return_logits=True, knn_method="brute")
net = B200RandLANet(9, 6, num_neighbors=16, return_logits=True)
net.load_state_dict(ref.state_dict(), strict=True)
net = net.to("cuda:0")
x, pos, y, batch, ptr = O.synthetic_batch([1500, 600], seed=7)
mask = (torch.rand(2100, 32) < 0.5).float() * 2.0
ref.train(), net.train()
ref.mlp_classif.injected_masks = [None, mask]
logits_ref = ref(x, pos, batch, ptr)
F.cross_entropy(logits_ref, y, ignore_index=65).backward()
net.injected_decimation_idx = ref.last_decimation_idx
net.mlp_classif.injected_masks = [None, mask.cuda()]
gref = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
lib = _lib.load()
saved = {}
for it, tc in enumerate((31, 15, 30, 17)):
    lib.b200_set_option(b"tensor_core_paths", tc)
    net.zero_grad(set_to_none=True)
    logits = net(x.cuda(), pos.cuda(), batch.cuda(), ptr.cuda())
    ops.cross_entropy(logits, y.cuda(), None, ignore_index=65).backward()
    rows = sorted(((float((p.grad.cpu() - gref[n]).norm() / (gref[n].norm() + 1e-30)), float(gref[n].norm()), n) for n, p in net.named_parameters() if p.grad is not None and "bias" not in n), reverse=True)
    print(f"run {it} tensor_cores={tc}: logits err {float((logits.detach().cpu() - logits_ref.detach()).abs().max()):.2e}")
    k = "fc0.weight"
    worst = max((float((p.grad.cpu() - gref[n]).norm() / (gref[n].norm() + 1e-30)), n) for n, p in net.named_parameters() if p.grad is not None and "lins.0.bias" not in n and "lins.1.bias" not in n)
    print(f"    fc0.weight rel {float((net.fc0.weight.grad.cpu() - gref[k]).norm() / gref[k].norm()):.2e}; worst {worst}")
    saved[tc] = {n: p.grad.cpu().clone() for n, p in net.named_parameters() if p.grad is not None}
k = "mlp_classif.norms.0.module.bias"
torch.set_printoptions(precision=3, linewidth=200, sci_mode=True)
print("ref ", gref[k][:16])
for tc in saved: print(tc, (saved[tc][k] - gref[k])[:64])
k = "mlp_classif.norms.0.module.weight"
for tc in saved: print(tc, (saved[tc][k] - gref[k])[:16])
