"""Diagnostic: per-stage divergence between B200RandLANet and the oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import randla_oracle as O
from tests.helpers import rand_cloud
from myria3d_b200 import B200RandLANet

def run(sizes, seed=0, train=True, dbl=False):
    torch.manual_seed(seed)
    ref = O.OracleRandLANet(9, 6, return_logits=True, knn_method="brute")
    g = torch.Generator().manual_seed(seed + 1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data.uniform_(0.7, 1.3, generator=g); m.bias.data.uniform_(-0.2, 0.2, generator=g)
            m.running_mean.uniform_(-0.2, 0.2, generator=g); m.running_var.uniform_(0.6, 1.4, generator=g)
    net = B200RandLANet(9, 6, return_logits=True); net.load_state_dict(ref.state_dict()); net.cuda()
    x, pos, b, ptr = rand_cloud(sizes, seed=seed)
    n = sum(sizes)
    ref.train(train); net.train(train)
    ref.mlp_classif.injected_masks = [None, torch.ones(n, 32)]
    net.mlp_classif.injected_masks = [None, torch.ones(n, 32).cuda()]
    out_ref = ref(x, pos, b, ptr).detach()
    refd = O.OracleRandLANet(9, 6, return_logits=True, knn_method="brute").double()
    refd.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in ref.state_dict().items()})
    refd.train(train); refd.mlp_classif.injected_masks = [None, torch.ones(n, 32).double()]
    # note: running stats of ref were updated by the first call; reload for fairness is not needed in train mode
    out_d = refd(x.double(), pos.double(), b, ptr, decimation_idx=ref.last_decimation_idx).detach()
    net.keep_stages = True
    net.injected_decimation_idx = ref.last_decimation_idx
    out = net(x.cuda(), pos.cuda(), b.cuda(), ptr.cuda()).detach().cpu()
    print(f"sizes={sizes} train={train}")
    for k in ref.stages:
        a, r, d = net.stages[k].detach().cpu().double(), ref.stages[k].detach().double(), refd.stages[k].detach()
        print(f"  {k:7s} shape {tuple(r.shape)}: |gpu-ref32| {float((a-r).abs().max()):.3e}  |gpu-ref64| {float((a-d).abs().max()):.3e}  |ref32-ref64| {float((r-d).abs().max()):.3e}  mag {float(d.abs().max()):.2e}")
    print(f"  logits: |gpu-ref32| {float((out-out_ref).abs().max()):.3e} |gpu-ref64| {float((out.double()-out_d).abs().max()):.3e} |ref32-ref64| {float((out_ref.double()-out_d).abs().max()):.3e}")

for sizes in ([50, 50], [700, 3, 40, 2000], [20, 20], [50, 50, 50, 50]):
    run(sizes, train=True)
run([50, 50], train=False)
