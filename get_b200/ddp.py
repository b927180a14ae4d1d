"""Data parallelism for the GET hot path: one process per GPU, claims sharded across ranks, ONE exchange
step per iteration -- a gradient all-reduce (NCCL over NVLink/NVSwitch) on a flat fp32 bucket holding only
the parameters that receive gradients. The reference is single-process
(no torch.distributed anywhere); this is the multi-GPU path BASELINE.json's north_star asks for.

Parameters without gradients in the reference (bilstm.*, query_bilstm.*, trans.*, ggnn_with_gsl.word_scorer1.*,
frozen embedding) are excluded, so no rank ever waits on them.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist

INERT_PREFIXES = ("bilstm.", "query_bilstm.", "trans.", "ggnn_with_gsl.word_scorer1.")


def trainable_named_parameters(model) -> List:
    return [(n, p) for n, p in model.named_parameters() if p.requires_grad and not n.startswith(INERT_PREFIXES)]


def shard_claims(evd_cnt, world_size: int):
    """Contiguous claim ranges per rank balanced by the number of evidences (cost ~ B1 = sum n_c).
    Returns a list of (start, stop) claim indices, one per rank (greedy prefix split)."""
    cnt = [int(c) for c in evd_cnt]
    total = sum(cnt)
    bounds, acc, start = [], 0, 0
    for r in range(world_size):
        target = total * (r + 1) / world_size
        stop = start
        while stop < len(cnt) and (acc + cnt[stop] <= target or stop == start) and len(cnt) - stop > world_size - r - 1:
            acc += cnt[stop]
            stop += 1
        if r == world_size - 1:
            stop = len(cnt)
        bounds.append((start, stop))
        start = stop
    return bounds


class FlatGradAllReduce(object):
    """Flat-bucket gradient averaging. `params` = the grad-receiving parameters in a fixed order."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None):
        self.params = list(params)
        self.group = process_group
        numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros((numel,), dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1

    def init_collective(self):
        """One all-reduce of the (zero) bucket: creates the communicator outside of any stream capture. Every rank must
        call it the same number of times."""
        if self.world > 1:
            dist.all_reduce(self.flat, group=self.group)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def reduce(self, weight: float = 1.0, collective: bool = True):
        """Average gradients over ranks (each rank's loss is a mean over its local claims; `weight` =
        local_claims * world / global_claims re-weights unequal shards). Leaves p.grad pointing into the bucket.
        collective=False only gathers the gradients into the bucket (warm-up steps that must not talk to other ranks)."""
        grads, views = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                grads.append(p.grad)
                views.append(v)
        if grads:
            torch._foreach_copy_(views, grads)
        if self.world > 1 and collective:
            if weight != 1.0:
                self.flat.mul_(weight)
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM,
                            group=self.group)
            if dist.get_backend(self.group) != "nccl":
                self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            p.grad = v
