"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: flat gradient all-reduce, BatchNorm-buffer broadcast
(torch DDP's broadcast_buffers=True), tile sharding."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from myria3d_b200.parallel import FlatGradAllReducer, broadcast_module_state, shard_tiles

    torch.manual_seed(100 + rank)  # different initial weights per rank ...
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.BatchNorm1d(8), torch.nn.Linear(8, 3))
    broadcast_module_state(net)  # ... made identical here
    w0 = net[0].weight.detach().clone()
    red = FlatGradAllReducer(net)
    assert red.check_views()
    x = torch.full((6, 4), float(rank + 1)) + torch.arange(6).float()[:, None]
    net(x).sum().backward()
    local = red.flat.clone()
    red.all_reduce()
    assert red.check_views()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    ok = torch.allclose(red.flat, expect, atol=1e-6) and torch.allclose(net[0].weight.grad.flatten(), expect[:32], atol=1e-6)
    ws = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    ok = ok and all(torch.equal(ws[0], w) for w in ws)
    red.zero_grad()
    ok = ok and float(net[2].bias.grad.abs().sum()) == 0.0
    ok = ok and shard_tiles(5, rank, world) == ([0, 2, 4] if rank == 0 else [1, 3, 0])
    # DDP's broadcast_buffers=True: after the (rank-dependent) forward above the BatchNorm running statistics differ;
    # broadcast_buffers() / its start + finish halves make every rank hold rank 0's, integer counters untouched
    rm_mine = net[1].running_mean.clone()
    rms = [torch.zeros_like(rm_mine) for _ in range(world)]
    dist.all_gather(rms, rm_mine)
    ok = ok and not torch.equal(rms[0], rms[1])
    nbt = int(net[1].num_batches_tracked)
    work = red.start_broadcast()
    red.finish_broadcast(work)
    ok = ok and torch.equal(net[1].running_mean, rms[0]) and int(net[1].num_batches_tracked) == nbt
    net[1].running_var.add_(float(rank))
    red.broadcast_buffers()
    rvs = [torch.zeros_like(rm_mine) for _ in range(world)]
    dist.all_gather(rvs, net[1].running_var.clone())
    ok = ok and torch.equal(rvs[0], rvs[1])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_shard_tiles_single_process():
    from myria3d_b200.parallel import shard_tiles

    assert shard_tiles(16, 3, 8) == [3, 11]
    assert shard_tiles(0, 0, 2) == []
    all_tiles = sorted(t for r in range(4) for t in shard_tiles(16, r, 4))
    assert all_tiles == list(range(16))
