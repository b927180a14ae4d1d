"""Host logic of the multi-GPU path on CPU: world_size-2 `gloo` run of the flat-bucket gradient all-reduce and
the claim sharding. The kernels themselves need a GPU; here the bucket logic is exercised with
plain CPU tensors standing in for gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from get_b200.ddp import FlatGradAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    red = FlatGradAllReduce(params)
    g = torch.Generator().manual_seed(100 + rank)
    params[0].grad = torch.randn(3, 4, generator=g)
    params[1].grad = torch.randn(5, generator=g)
    params[2].grad = None                       # a parameter without a gradient on this rank
    red.reduce()
    ok = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, red.views))
    out.put((rank, [p.grad.clone() for p in params], ok, red.nbytes))
    # second iteration: fresh grads replace the bucket views, as after optimizer.zero_grad(set_to_none=True)
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.zeros(5)
    params[2].grad = torch.ones(2, 2) * (rank + 1)
    red.reduce()
    out.put((rank, [p.grad.clone() for p in params], True, red.nbytes))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    first = {r: g for r, g, ok, nb in res[:] if True}
    by_rank = {0: [], 1: []}
    for r, g, ok, nb in res:
        assert ok and nb == (12 + 5 + 4) * 4
        by_rank[r].append(g)
    g0 = torch.Generator().manual_seed(100)
    g1 = torch.Generator().manual_seed(101)
    a0, b0 = torch.randn(3, 4, generator=g0), torch.randn(5, generator=g0)
    a1, b1 = torch.randn(3, 4, generator=g1), torch.randn(5, generator=g1)
    for r in (0, 1):
        it1, it2 = by_rank[r]
        assert torch.allclose(it1[0], (a0 + a1) / 2) and torch.allclose(it1[1], (b0 + b1) / 2)
        assert torch.equal(it1[2], torch.zeros(2, 2))
        assert torch.allclose(it2[0], torch.full((3, 4), 1.5)) and torch.allclose(it2[2], torch.full((2, 2), 1.5))


def test_shard_claims_balances_by_evidence_count():
    from get_b200.ddp import shard_claims
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        cnt = np.clip(rng.geometric(1 / 6.74, size=32), 1, 30)
        bounds = shard_claims(cnt, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == 32
        assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
        assert all(hi > lo for lo, hi in bounds)
        loads = [int(cnt[lo:hi].sum()) for lo, hi in bounds]
        assert max(loads) <= cnt.sum() / world + 30
    assert shard_claims([5, 5], 2) == [(0, 1), (1, 2)]


def test_trainable_parameters_exclude_inert_ones():
    from get_b200 import synthetic
    from get_b200.ddp import trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from helpers import load_golden
    model = Graph_basedSemantiStructure(synthetic.match_params(synthetic.get_workload("tiny")))
    names = [n for n, _ in trainable_named_parameters(model)]
    gold = load_golden("tiny_snopes")
    with_grad = sorted(k[5:] for k in gold if k.startswith("grad/"))
    assert sorted(names) == with_grad     # exactly the parameters that get a gradient in the reference


def _worker_local_gather(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from get_b200.ddp import FlatGradAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
    red = FlatGradAllReduce(params)
    red.init_collective()                      # symmetric, outside any step
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    red.reduce(collective=False)               # warm-up semantics: gather only, no communication
    local = red.flat.clone()
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    red.reduce()                               # the real step: averaged over ranks
    q.put((rank, local.tolist(), red.flat.tolist(), all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, red.views))))
    dist.destroy_process_group()


def test_reduce_without_collective_only_gathers_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_local_gather, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, local, reduced, aliased in res:
        assert local == [float(rank + 1)] * 11
        assert reduced == [1.5] * 11
        assert aliased
