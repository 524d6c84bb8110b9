"""Whole-step kernel breakdown with torch.profiler (CUPTI): every kernel, ours and torch's.
Run on the GPU box: python scripts/profile_step.py [--graphed] > gpurun_out/profile_step.txt"""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

class A: pass
args = A(); args.tiles, args.points = 16, 12800
from myria3d_b200 import Model
from myria3d_b200.parallel import FlatGradAllReducer
from myria3d_b200.graphed import GraphedTrainStep
dev = torch.device("cuda", 0)
torch.manual_seed(12345)
model = Model(neural_net_class_name="B200RandLANet", neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
              criterion=torch.nn.CrossEntropyLoss(ignore_index=65), lr=bench.LR).to(dev).train()
red = FlatGradAllReducer(model)
graphed = "--graphed" in sys.argv
from myria3d_b200.optim import FlatAdam
opt = FlatAdam(model, lr=bench.LR, reducer=red)
b = bench.host_batch(16, 12800, 12345).to(dev)
model.model.decimation_rng = "fused"
step_g = GraphedTrainStep(model, opt, red) if graphed else None
def step():
    if graphed:
        return step_g(b)
    red.zero_grad(); out = model.training_step(b, 0); out["loss"].backward(); opt.step(); return out["loss"]
for _ in range(3): step()
torch.cuda.synchronize()
NSTEP = 3
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NSTEP): step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name).replace("at::native::", "").replace("b200::", "")[:100]
        agg[name][0] += 1; agg[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if re.search(r"(lfa_|knn_|linear_|affine_|bn_finalize|_rows_|interp|moments|fold)", k))
print(f"{'graphed' if graphed else 'eager'}: {NSTEP} steps, device kernel time {tot/NSTEP/1e3:.3f} ms/step, libb200randla {ours/NSTEP/1e3:.3f} ms/step, other {(tot-ours)/NSTEP/1e3:.3f} ms/step")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{us/NSTEP:9.1f} us/step {100*us/tot:5.1f}%  n/step={n/NSTEP:6.1f}  {k}")
