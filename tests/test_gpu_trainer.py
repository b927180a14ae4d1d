"""Thin data-parallel trainer (get_b200/trainer.py), flat Adam, gradient sinks and the multi-GPU step (`-m gpu`)."""
import os
import socket

import numpy as np
import pytest
import torch

from get_b200 import synthetic
from get_b200.keywords import KeyWordSettings as K

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(w, seed=4, train=True):
    from get_b200.model import Graph_basedSemantiStructure
    torch.manual_seed(seed)
    m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(DEV)
    m.train(train)
    return m


def test_flat_adam_matches_torch_adam_and_sinks_match_autograd():
    """Eval mode (no dropout) so both runs see the same function: (a) gradients written straight into the flat bucket equal
    the autograd gradients bit for bit; (b) the one-kernel flat Adam follows torch.optim.Adam(weight_decay)."""
    from get_b200 import ops
    from get_b200.ddp import FlatAdam, FlatGradAllReduce, trainable_named_parameters
    w = synthetic.get_workload("snopes", batch_claims=4, vocab=300, n_article_sources=8)
    tens = [synthetic.batch_to_torch(synthetic.make_batch(w, seed=s), device=DEV) for s in (1, 2)]

    def run(flat):
        m = _model(w, train=False)
        named = trainable_named_parameters(m)
        params = [p for _, p in named]
        if flat:
            red = FlatGradAllReduce(params, names=[n for n, _ in named]).attach()
            opt = FlatAdam(red, lr=1e-3, weight_decay=1e-3)
        else:
            opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-3)
        losses, grads0 = [], None
        for it in range(5):
            q, d, l, kw = tens[it % 2]
            if flat:
                red.zero()
            else:
                opt.zero_grad(set_to_none=True)
            loss = ops.cross_entropy(m(q, d, **kw), l)
            loss.backward()
            if it == 0:
                grads0 = {n: p.grad.clone() for n, p in named}
            opt.step()
            losses.append(float(loss))
        if flat:
            red.detach()
        return losses, grads0, {n: p.detach().clone() for n, p in named}

    l1, g1, p1 = run(False)
    l2, g2, p2 = run(True)
    for n in g1:
        assert torch.equal(g1[n], g2[n]), "sink gradient differs from the autograd gradient: %s" % n
    assert np.allclose(l1, l2, rtol=0, atol=2e-6), (l1, l2)
    for n in p1:
        assert float((p1[n] - p2[n]).abs().max()) < 5e-6, n      # 5 steps of lr 1e-3: 1e-3 of the distance travelled


def test_eager_forward_after_replays_sees_the_current_weights():
    """a replayed step changes the weights on the device; an eager forward issued afterwards (evaluation) must
    not read packed weights of an older step. Interleave replays and eager evaluations against the exact SIMT path."""
    from get_b200 import ops
    from get_b200.ddp import FlatAdam, FlatGradAllReduce, trainable_named_parameters
    from get_b200.step_graph import CapturedTrainStep, pad_batch
    w = synthetic.get_workload("snopes", batch_claims=4, vocab=300, n_article_sources=8)
    batch = pad_batch(synthetic.make_batch(w, seed=3), 32)
    q, d, l, kw = synthetic.batch_to_torch(batch, device=DEV)
    m = _model(w, train=False)
    named = trainable_named_parameters(m)
    red = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
    opt = FlatAdam(red, lr=5e-3, weight_decay=1e-3)
    step = CapturedTrainStep(m, opt, red)
    for rnd in range(3):
        for _ in range(2):
            step.step(q, d, l, kw, batch["n_real_claims"])
        with torch.no_grad():
            tc = m(q, d, **kw).clone()
            ops.TC_ENABLED = False
            try:
                ref = m(q, d, **kw).clone()
            finally:
                ops.TC_ENABLED = True
        assert float((tc - ref).abs().max()) < 1e-4, (rnd, float((tc - ref).abs().max()))
    red.detach()


def test_trainer_fits_evaluates_and_checkpoints(tmp_path):
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.trainer import GETTrainer
    w = synthetic.get_workload("snopes", batch_claims=6, vocab=300, n_article_sources=8)
    train = [synthetic.make_batch(w, seed=10 + i) for i in range(3)]
    val = [synthetic.make_batch(w, seed=50)]
    m = _model(w)
    path = str(tmp_path / "Fold_0" / "saved_model_1")
    tr = GETTrainer(m, lr=1e-3, n_iter=3, saved_model=path, pad_pairs_to=32)
    out = tr.fit(lambda epoch: train, val)
    assert len(out["history"]) == 3 and all(np.isfinite(h["epoch_loss"]) and h["steps"] == 3 for h in out["history"])
    assert out["history"][-1]["epoch_loss"] < out["history"][0]["epoch_loss"], "training must reduce the loss on 3 fixed batches"
    assert "val_f1_macro" in out["history"][0] and "val_auc" in out["history"][0]
    if out["best_val_f1_macro"] > 0:
        sd = torch.load(path, map_location="cpu")
        ref_keys = set(Graph_basedSemantiStructure(synthetic.match_params(w)).state_dict().keys())
        assert set(sd.keys()) == ref_keys                      # reference key names: the reference's load_best_model reads it
        tr.load_best_model()
    tr.reducer.detach()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    from get_b200 import ops
    from get_b200.ddp import FlatGradAllReduce, shard_claims, trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep, pad_batch, slice_batch
    w = synthetic.get_workload("snopes", batch_claims=8, vocab=300, n_article_sources=8)
    batch = synthetic.make_batch(w, seed=7)
    torch.manual_seed(4)
    m = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev).eval()
    named = trainable_named_parameters(m)
    red = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
    step = CapturedTrainStep(m, None, red)                    # no optimizer: the bucket keeps the all-reduced gradients
    lo, hi = shard_claims(batch[K.EvidenceCountPerQuery], world)[rank]
    local = pad_batch(slice_batch(batch, lo, hi), 16)
    t = synthetic.batch_to_torch(local, device=dev)
    B = batch["query"].shape[0]
    step.step(*t, local["n_real_claims"], global_claims=B)
    step.step(*t, local["n_real_claims"], global_claims=B)    # second call = pure graph replay (overlapped chunk all-reduces)
    torch.cuda.synchronize()
    got = {n: p.grad.detach().cpu().clone() for n, p in named}
    if rank == 0:
        # single-GPU gradient of the mean loss over the WHOLE batch
        red.detach()
        m.zero_grad(set_to_none=True)
        q_, d_, l_, kw_ = synthetic.batch_to_torch(batch, device=dev)
        ops.cross_entropy(m(q_, d_, **kw_), l_).backward()
        ref = {n: p.grad.detach().cpu().clone() for n, p in named}
        err = max(float((got[n] - ref[n]).abs().max()) for n in ref)
        print("2-GPU vs 1-GPU gradient max abs diff: %.3e" % err, flush=True)
        q.put(err)
        q.close()
        q.join_thread()           # the result is flushed to the parent before this process leaves through os._exit
    dist.barrier()
    torch.cuda.synchronize()
    # captured graphs hold NCCL kernels: tearing the communicator down under them can block; leave like bench.py does
    os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_gradients_equal_single_gpu_gradients():
    """gradients after the (chunked, overlapped, shard-weighted) all-reduce over 2 GPUs == the
    1-GPU gradients on the concatenated batch."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6, err
