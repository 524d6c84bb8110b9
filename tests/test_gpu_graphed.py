"""CUDA-graph step (GraphedTrainStep) vs the eager step: same losses, parameters move identically."""
import copy

import pytest
import torch

from oracle import randla_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(seed=0):
    from myria3d_b200 import Model

    torch.manual_seed(seed)
    return Model(neural_net_class_name="B200RandLANet",
                 neural_net_hparams=dict(num_features=9, num_classes=6, num_neighbors=16, decimation=4, return_logits=True),
                 criterion=torch.nn.CrossEntropyLoss(ignore_index=65), lr=1e-3).to(DEV).train()


def _batch(sizes, seed):
    from myria3d_b200 import Batch, Data

    datas = []
    for i, n in enumerate(sizes):
        x, pos, y = O.synthetic_tile(n, seed + i)
        datas.append(Data(x=x, pos=pos, y=y))
    return Batch.from_data_list(datas)


def test_graphed_forward_matches_eager(lib):
    """The loss a replay reports == the eager training_step loss at the same parameters, same decimation
    subsets (dropout disabled so the two Philox consumers cannot differ)."""
    from myria3d_b200.graphed import GraphedTrainStep
    from myria3d_b200.parallel import FlatGradAllReducer

    m = _model(1)
    m.model.mlp_classif.dropout = [0.0, 0.0]
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
    step = GraphedTrainStep(m, opt, FlatGradAllReducer(m))
    b = _batch([900, 400], 3).pin_memory()
    step(b)  # capture (3 eager warm-up steps) + first replay
    key = tuple(b.ptr.tolist())
    for _ in range(3):
        snap = copy.deepcopy(m)  # parameters before the replay
        loss_g = float(step(b))
        cap = step._captured[key]
        snap.model.injected_decimation_idx = [t.clone() for t in cap.idx_static]
        out = snap.training_step(b.to(DEV), 0)
        assert abs(float(out["loss"].detach()) - loss_g) < 1e-4, (float(out["loss"].detach()), loss_g)
    assert step.library_launches >= 4 * 200


def test_graphed_step_trains(lib):
    """Loss decreases over replays on a fixed batch; one graph per layout; e2e from pinned host memory."""
    from myria3d_b200.graphed import GraphedTrainStep

    m = _model(2)
    opt = torch.optim.Adam(m.parameters(), lr=3e-3, capturable=True)
    step = GraphedTrainStep(m, opt)
    b = _batch([700, 700], 5).pin_memory()
    first = float(step(b))
    for _ in range(30):
        last = float(step(b))
    assert last < first, (first, last)
    assert len(step._captured) == 1
    b2 = _batch([500, 300, 200], 9)
    step(b2.to(DEV))
    assert len(step._captured) == 2
    assert torch.isfinite(step.last_outputs(b2)["logits"]).all()


def test_flat_adam_matches_torch_adam(lib):
    """b200_adam_flat == torch.optim.Adam on the same gradients over several steps (parameters as flat views)."""
    from myria3d_b200.optim import FlatAdam

    torch.manual_seed(0)
    # (no BatchNorm right after a Linear: that Linear's bias has a mathematically zero gradient, whose fp32 noise Adam
    # normalises to +-lr steps -- any two Adam implementations diverge there by O(lr))
    net_a = torch.nn.Sequential(torch.nn.Linear(9, 33), torch.nn.Tanh(), torch.nn.Linear(33, 7)).to(DEV)
    net_b = copy.deepcopy(net_a)
    opt_a = torch.optim.Adam(net_a.parameters(), lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    opt_b = FlatAdam(net_b, lr=3e-3)
    for s in range(5):
        x = torch.randn(64, 9, device=DEV)
        for net, opt in ((net_a, opt_a), (net_b, opt_b)):
            opt.zero_grad()
            net(x).square().mean().backward()
            opt.step()
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6, rtol=1e-5), float((pa - pb).abs().max())
    assert int(opt_b.step_count) == 5
