"""CPU, world_size 2 over gloo: the gradient exchange of the data-parallel training step and the lattice sharding
rule used for multi-GPU panorama generation."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spgan_b200.training import allreduce_gradients, replicas_in_sync, sync_module_states
    from spgan_b200 import panorama
    # replicas: different init per rank -> broadcast from rank 0 -> identical; a data-parallel SGD loop on rank-dependent
    # data keeps them identical (the ranks must not train different models)
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 2))
    differs_before = not replicas_in_sync([net], world)
    nb = sync_module_states([net], world)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    gdata = torch.Generator().manual_seed(1000 + rank)
    for _ in range(3):
        opt.zero_grad()
        net(torch.randn(8, 6, generator=gdata)).square().mean().backward()
        allreduce_gradients(net.parameters(), world, bucket_bytes=64)
        opt.step()
    dp_ok = differs_before and nb == 9 and replicas_in_sync([net], world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (3, 1000, 70000, 5)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params.append(torch.nn.Parameter(torch.zeros(4)))  # no grad: must be skipped
    n = allreduce_gradients(params, world, bucket_bytes=1 << 12)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params[:4]))
    # lattice sharding: every position owned by exactly one rank
    pl = panorama.plan(384, 768)
    pos = panorama.positions(pl)
    mine = pos[rank::world]
    counts = torch.zeros(len(pos))
    for p_ in mine:
        counts[pos.index(p_)] += 1
    dist.all_reduce(counts)
    # sharded panorama generation with a stub generator: identical to the sequential loop on every rank
    def stub_gen(global_latent, lat, coords, cp, noises=None, styles=None):
        v = lat.mean(dim=(1, 2, 3)) + coords.mean(dim=(1, 2, 3)) + noises[7].mean(dim=(1, 2, 3)) + cp["p_x_st"] + 3 * cp["p_y_st"]
        base = torch.arange(101 * 101, dtype=torch.float32).view(1, 1, 101, 101) / 1e4
        return (v.view(-1, 1, 1, 1) + base).expand(-1, 3, -1, -1).contiguous()
    g = torch.Generator().manual_seed(5)
    gl = torch.randn(2, 512, generator=g)
    canvas = torch.randn(2, 256, pl["lat_h"], pl["lat_w"], generator=g)
    noises = [torch.randn(2, 1, pl["noise_h"][l], pl["noise_w"][l], generator=g) for l in range(8)]
    seq = panorama.generate(stub_gen, pl, gl, canvas, noises)
    sh = panorama.generate_sharded(stub_gen, pl, gl, canvas, noises, rank, world)
    same = bool(torch.equal(seq, sh))
    ret[rank] = bool(dp_ok and ok and n >= 2 and params[4].grad is None and bool((counts == 1).all()) and same)
    dist.destroy_process_group()


def test_gradient_allreduce_and_lattice_sharding_world2():
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
