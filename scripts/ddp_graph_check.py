"""torchrun --nproc-per-node 2 scripts/ddp_graph_check.py : losses / parameter and buffer checksums of 6 graphed steps,
with the collectives captured in the step graph (default) or issued eagerly between two graphs (B200_COLLECTIVES_IN_GRAPH=0).
Both modes must print the same numbers (to fp32 atomics noise) and every rank the same buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from myria3d_b200 import Model
from myria3d_b200.parallel import FlatGradAllReducer, broadcast_module_state
from myria3d_b200.graphed import GraphedTrainStep
from myria3d_b200.optim import FlatAdam

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.manual_seed(12345)
model = Model(neural_net_class_name="B200RandLANet", neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
              criterion=torch.nn.CrossEntropyLoss(ignore_index=65), lr=bench.LR).to(dev).train()
broadcast_module_state(model)
red = FlatGradAllReducer(model)
opt = FlatAdam(model, lr=bench.LR, reducer=red)
step = GraphedTrainStep(model, opt, red)
b = bench.host_batch(4, 12800, 100 + rank).to(dev)
losses = []
for i in range(6):
    losses.append(float(step(b)))
torch.cuda.synchronize()
psum = float(sum(p.double().abs().sum() for p in model.parameters()))
bsum = float(sum(bf.double().abs().sum() for bf in model.buffers() if bf.is_floating_point()))
out = [None] * world
dist.all_gather_object(out, (rank, losses, psum, bsum))
if rank == 0:
    print("collectives_in_graph", step.collectives_in_graph)
    for o in out: print(o)
dist.destroy_process_group()
