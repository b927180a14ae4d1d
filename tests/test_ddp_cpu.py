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
        assert ok and nb == 32 * 4          # 21 floats, bucket padded to a multiple of 32
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
    q.put((rank, local[:11].tolist(), red.flat[:11].tolist(), all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, red.views))))
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


def _worker_attached(rank, world, port, q):
    """Attached bucket (gradient storage) + unequal shards: the weighted average must equal the global-batch gradient."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from get_b200 import ops
    from get_b200.ddp import FlatGradAllReduce
    names = ["out.0.weight", "ggnn_with_gsl.feat_prop1.proj.linear.weight", "ggnn_with_gsl.feat_prop2.proj.linear.weight",
             "self_att_word.linear1.weight"]
    params = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3, 2)),
              torch.nn.Parameter(torch.zeros(5))]
    red = FlatGradAllReduce(params, names=names).attach()
    # bucket order = backward completion order: head chunk, feat_prop2, feat_prop1
    assert red.names == ["out.0.weight", "self_att_word.linear1.weight", "ggnn_with_gsl.feat_prop2.proj.linear.weight",
                         "ggnn_with_gsl.feat_prop1.proj.linear.weight"]
    assert red.chunks == [(0, 11), (32, 38), (64, 68)]          # chunks start on 32-float (128-byte) boundaries
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    assert all(ops.sink_of(p) is not None for p in red.params)
    n_local, n_global = (3, 8) if rank == 0 else (5, 8)          # claims on this rank / in the global batch
    red.zero()
    for i, p in enumerate(red.params):                           # "local mean" gradients: value = rank + 1 everywhere
        p.grad.add_(float(rank + 1))
    red.set_weight(n_local * world / n_global)
    red.reduce()
    want = (3 * 1.0 + 5 * 2.0) / 8                               # gradient of the global-batch mean
    ok = all(torch.allclose(p.grad, torch.full_like(p.grad, want)) for p in red.params)
    # dropping p.grad (zero_grad(set_to_none=True)) disables the sink for that parameter: plain autograd flow
    red.params[0].grad = None
    ok = ok and ops.sink_of(red.params[0]) is None and ops.sink_of(red.params[1]) is not None
    red.detach()
    ok = ok and ops.sink_of(red.params[1]) is None
    q.put((rank, ok))
    dist.destroy_process_group()


def test_attached_bucket_chunks_and_shard_weights_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_attached, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_slice_batch_partitions_a_global_batch():
    from get_b200 import synthetic
    from get_b200.ddp import shard_claims
    from get_b200.keywords import KeyWordSettings as K
    from get_b200.step_graph import pad_batch, slice_batch
    w = synthetic.get_workload("tiny", batch_claims=9)
    b = synthetic.make_batch(w, seed=5)
    for world in (2, 4):
        parts = [slice_batch(b, lo, hi) for lo, hi in shard_claims(b[K.EvidenceCountPerQuery], world)]
        assert sum(p["pairs"] for p in parts) == b["pairs"]
        for key in (K.Evd_Docs_Adj, K.DocContentNoPaddingEvidence, "e_lens"):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), b[key])
        for key in ("query", "labels", K.Query_Adj, K.DocSources):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), b[key])
        for p in parts:
            pp = pad_batch(p, 8)
            assert pp["pairs"] % 8 == 0 and pp["n_real_claims"] == p["query"].shape[0]


def test_classification_metrics_match_sklearn():
    from get_b200.trainer import classification_metrics
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    y = rng.integers(0, 2, size=200)
    s = rng.normal(size=200) + y
    p = (s > 0.4).astype(int)
    m = classification_metrics(y, p, s)
    fpr, tpr, _ = sk.roc_curve(y, s, pos_label=1)
    assert abs(m["auc"] - sk.auc(fpr, tpr)) < 1e-12
    assert abs(m["f1_macro"] - sk.f1_score(y, p, average="macro")) < 1e-12
    assert abs(m["f1_micro"] - sk.f1_score(y, p, average="micro")) < 1e-12
    assert abs(m["precision_true_cls"] - sk.precision_score(y, p, labels=[1], average=None)[0]) < 1e-12
    assert abs(m["recall_false_cls"] - sk.recall_score(y, p, labels=[0], average=None)[0]) < 1e-12


def test_balanced_claim_assignment_and_selection():
    """balance_claims: a partition of the claims whose per-rank evidence totals differ by less than the largest claim of the
    lightest tail; select_claims gathers exactly those claims' rows."""
    from get_b200 import synthetic
    from get_b200.ddp import balance_claims
    from get_b200.keywords import KeyWordSettings as K
    from get_b200.step_graph import select_claims
    rng = np.random.default_rng(3)
    for world in (2, 4, 8):
        cnt = rng.integers(1, 30, size=32 * world)
        parts = balance_claims(cnt, world)
        assert sorted(i for p in parts for i in p) == list(range(len(cnt)))
        loads = [int(cnt[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= 2, loads
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 6
    w = synthetic.get_workload("tiny", batch_claims=9)
    b = synthetic.make_batch(w, seed=5)
    parts = balance_claims(b[K.EvidenceCountPerQuery], 2)
    subs = [select_claims(b, p) for p in parts]
    assert sum(s["pairs"] for s in subs) == b["pairs"]
    off = np.concatenate([[0], np.cumsum(b[K.EvidenceCountPerQuery])])
    for p, s in zip(parts, subs):
        assert np.array_equal(s["labels"], b["labels"][p])
        rows = np.concatenate([np.arange(off[c], off[c + 1]) for c in p])
        assert np.array_equal(s[K.Evd_Docs_Adj], b[K.Evd_Docs_Adj][rows])
        assert np.array_equal(s[K.EvidenceCountPerQuery], np.asarray(b[K.EvidenceCountPerQuery])[p])
