"""Which aten ops (torch-side glue) run in one eager step, with counts and shapes."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from myria3d_b200 import Model
from myria3d_b200.parallel import FlatGradAllReducer
dev = torch.device("cuda", 0)
model = Model(neural_net_class_name="B200RandLANet", neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
              criterion=torch.nn.CrossEntropyLoss(ignore_index=65), lr=bench.LR).to(dev).train()
model.model.decimation_rng = "fused"
red = FlatGradAllReducer(model)
opt = torch.optim.Adam(model.parameters(), lr=bench.LR, capturable=True, fused=True)
b = bench.host_batch(16, 12800, 12345).to(dev)
def step():
    red.zero_grad(); out = model.training_step(b, 0); out["loss"].backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
if "--capture" in sys.argv:
    from myria3d_b200.graphed import GraphedTrainStep
    sg = GraphedTrainStep(model, opt, red, warmup_steps=1)
    with profile(activities=[ProfilerActivity.CPU], record_shapes=True) as prof:
        sg._captured[tuple(b.ptr.tolist())] = sg._capture(b)
    torch.cuda.synchronize()
else:
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.count > 0:
        rows.append(((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total) or e.cpu_time_total, e.count, e.key, str(e.input_shapes)[:90]))
rows.sort(reverse=True)
agg = collections.defaultdict(lambda: [0, 0.0])
for t, c, k, s in rows:
    agg[k][0] += c; agg[k][1] += t
print("per-op totals:")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {t:9.1f} us  n={c:5d}  {k}")
print("top (op, shapes):")
for t, c, k, s in rows[:40]:
    print(f"  {t:9.1f} us  n={c:4d}  {k:28s} {s}")
