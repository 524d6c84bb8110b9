"""Throughput of the two non-headline BASELINE.json configurations on one B200 (CUDA events, resident inputs):
  D  configs[3]: predict path -- eval forward on 40 960-point tiles (batch 50), k=10 interpolation of the logits to every
     point of the 60 000-point windows, sliding-window stitch (scatter-sum, softmax, argmax, entropy) of the batch;
  E  configs[4]: K=32, 65 536-point tiles, batch 4 -- train step fwd + CE + bwd (eager, torch Adam excluded).
Prints one JSON line per configuration (not the bench.py contract: these are profile numbers)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from myria3d_b200 import Batch, Data, Model, ops
from myria3d_b200.interpolation import Interpolator
from myria3d_b200.synthetic import synthetic_tile

dev = torch.device("cuda", 0)
CLASSES = {1: "unclassified", 2: "ground", 6: "building", 9: "water", 17: "bridge", 64: "lasting_above"}


def timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]


def config_d(tiles=50, full=60000, sub=40960):
    model = Model(neural_net_class_name="B200RandLANet",
                  neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
                  criterion=torch.nn.CrossEntropyLoss(ignore_index=65), interpolation_k=10, num_workers=1).to(dev).eval()
    g = torch.Generator().manual_seed(0)
    datas = []
    for w in range(tiles):
        x, pos, y = synthetic_tile(full, seed=900 + w)
        keep = torch.randperm(full, generator=g)[:sub]
        d = Data(x=x[keep], pos=pos[keep], y=y[keep])
        d.copies = {"pos_copy": pos, "pos_sampled_copy": pos[keep]}
        d.idx_in_original_cloud = np.arange(w * 45000, w * 45000 + full, dtype=np.int64)
        datas.append(d)
    batch = Batch.from_data_list(datas).to(dev)
    nb_points = 45000 * (tiles - 1) + full

    def step():
        with torch.no_grad():
            _, logits = model.forward(batch)  # network + k=10 interpolation (models/model.py:86-98), stays on the GPU
            itp = Interpolator(interpolation_k=10, classification_dict=CLASSES)
            itp.store_predictions(logits, batch.idx_in_original_cloud)
            reduced, idx, _ = itp._reduce(nb_points)
            return ops.stitch_finalize(reduced, idx, want_logits=False)

    ms = timed(step)
    print(json.dumps({"config": "D: predict path, 50 x 40960-pt tiles -> 50 x 60000-pt windows, k=10 interpolation + stitch",
                      "ms_per_batch": ms, "subsampled_points_per_s": tiles * sub / ms * 1e3,
                      "full_cloud_points_per_s": tiles * full / ms * 1e3}), flush=True)


def config_e(tiles=4, pts=65536):
    from myria3d_b200 import B200RandLANet
    torch.manual_seed(0)
    net = B200RandLANet(9, 6, num_neighbors=32, return_logits=True).to(dev).train()
    xs = [synthetic_tile(pts, seed=70 + t) for t in range(tiles)]
    x = torch.cat([a[0] for a in xs]).to(dev); pos = torch.cat([a[1] for a in xs]).to(dev); y = torch.cat([a[2] for a in xs]).to(dev)
    batch = torch.arange(tiles).repeat_interleave(pts).to(dev)
    ptr = torch.arange(tiles + 1, dtype=torch.int64, device=dev) * pts
    net.decimation_rng = "fused"

    def step():
        for p in net.parameters():
            p.grad = None
        loss = ops.cross_entropy(net(x, pos, batch, ptr), y, None, 65)
        loss.backward()
        return loss

    ms = timed(step)
    print(json.dumps({"config": "E: K=32, 4 x 65536-pt tiles, train fwd + CE + bwd (eager launches)", "ms_per_step": ms,
                      "points_per_s": tiles * pts / ms * 1e3}), flush=True)


if __name__ == "__main__":
    config_e()
    config_d()
