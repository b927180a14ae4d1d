"""Data parallelism for the GET hot path: one process per GPU, claims sharded across ranks, ONE exchange
step per iteration -- a gradient all-reduce (NCCL over NVLink/NVSwitch) on a flat fp32 bucket holding only
the parameters that receive gradients. The reference is single-process
(no torch.distributed anywhere); this is the multi-GPU path BASELINE.json's north_star asks for.

Parameters without gradients in the reference (bilstm.*, query_bilstm.*, trans.*, ggnn_with_gsl.word_scorer1.*,
frozen embedding) are excluded, so no rank ever waits on them.

The bucket is laid out in the order the backward pass finishes with the parameters -- [attention / MLP / source
embeddings | feat_prop2 | feat_prop1 + claim GGNN] -- and, once `attach()`ed, it IS the gradient storage: the weight
gradient kernels accumulate straight into it (get_b200.ops.GRAD_SINK), every p.grad is a view of it, and each chunk is
all-reduced on a communication stream as soon as the backward pass has left its layers (get_b200.ops.grad_marker), so
only the last chunk's collective is exposed.
"""
import weakref
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

INERT_PREFIXES = ("bilstm.", "query_bilstm.", "trans.", "ggnn_with_gsl.word_scorer1.")
# reduce order = order in which the backward pass completes the gradients
CHUNK_OF_PREFIX = (("ggnn_with_gsl.feat_prop2.", 1), ("ggnn_with_gsl.feat_prop1.", 2), ("ggnn4claim_1.", 2))
N_CHUNKS = 3
CHUNK_TAGS = {"head": 0, "feat_prop2": 1}      # ops.grad_marker tag -> chunk that is complete when it fires


def trainable_named_parameters(model) -> List:
    return [(n, p) for n, p in model.named_parameters() if p.requires_grad and not n.startswith(INERT_PREFIXES)]


def chunk_of(name: str) -> int:
    for prefix, c in CHUNK_OF_PREFIX:
        if name.startswith(prefix):
            return c
    return 0


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


def balance_claims(evd_cnt, world_size: int):
    """Claims per rank (lists of claim indices, each sorted) balanced by evidence count AND claim count: claims in order
    of decreasing evidence count go to the rank with the fewest evidences so far (ties: fewest claims, lowest rank). The
    per-rank evidence totals then differ by at most one claim's worth of the SMALLEST claims, typically 0-1 pairs, where
    a contiguous split (shard_claims) is off by up to one whole claim."""
    cnt = [int(c) for c in evd_cnt]
    order = sorted(range(len(cnt)), key=lambda i: (-cnt[i], i))
    load, parts = [0] * world_size, [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], len(parts[k]), k))
        parts[r].append(i)
        load[r] += cnt[i]
    return [sorted(p) for p in parts]


class FlatGradAllReduce(object):
    """Flat-bucket gradient averaging. `params` = the grad-receiving parameters; `names` (optional, same order) places
    them in backward-completion order so that the bucket can be reduced in chunks that overlap the backward pass."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, names: Optional[Sequence[str]] = None):
        params = list(params)
        if names is not None:
            names = list(names)
            order = sorted(range(len(params)), key=lambda i: (chunk_of(names[i]), i))
            params = [params[i] for i in order]
            names = [names[i] for i in order]
        self.params, self.names = params, names
        self.group = process_group
        # chunks start on 128-byte boundaries (collectives and vectorised kernels on a chunk see an aligned buffer); the
        # padding floats stay zero for ever: zero gradient, zero parameter, zero Adam update
        ALIGN = 32
        offs, off, prev = [], 0, None
        for i, p in enumerate(self.params):
            c = chunk_of(names[i]) if names is not None else 0
            if prev is not None and c != prev:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
            offs.append(off)
            off += p.numel()
            prev = c
        numel = (off + ALIGN - 1) // ALIGN * ALIGN
        dev = self.params[0].device
        self.flat = torch.zeros((numel,), dtype=torch.float32, device=dev)
        self.views, self.bounds, self.offsets = [], [], offs
        chunk_lo = [None] * N_CHUNKS
        chunk_hi = [0] * N_CHUNKS
        for i, p in enumerate(self.params):
            off = offs[i]
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            c = chunk_of(names[i]) if names is not None else 0
            if chunk_lo[c] is None:
                chunk_lo[c] = off
            chunk_hi[c] = off + p.numel()
        self.chunks = [(lo, hi) for lo, hi in zip(chunk_lo, chunk_hi) if lo is not None]
        self._chunk_index = {c: k for k, c in enumerate(c for c in range(N_CHUNKS) if chunk_lo[c] is not None)}
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.attached = False
        self._comm_stream = None
        self._pending = []          # chunks already handed to the communication stream in this step
        self._weight = 1.0
        self.overlap = True         # chunk all-reduces from the backward pass (False: one all-reduce in reduce())

    # ---------------------------------------------------------------------------------------------------------
    def init_collective(self):
        """One all-reduce of the (zero) bucket: creates the communicator outside of any stream capture. Every rank must
        call it the same number of times."""
        if self.world > 1:
            dist.all_reduce(self.flat, group=self.group)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def _avg_op(self):
        return dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM

    def _all_reduce(self, t: torch.Tensor):
        if self._weight != 1.0:
            t.mul_(self._weight)
        dist.all_reduce(t, op=self._avg_op(), group=self.group)
        if dist.get_backend(self.group) != "nccl":
            t.div_(self.world)

    # ---------------------------------------------------------------------------------------------------------
    def attach(self):
        """Make the bucket THE gradient storage: p.grad = view (autograd accumulates in place), the weight-gradient kernels
        of get_b200.ops write into the views directly, and chunk all-reduces are issued from the backward pass. The owner
        calls zero() at the start of every step instead of optimizer.zero_grad()."""
        from . import ops
        for p, v in zip(self.params, self.views):
            p.grad = v
            ops.GRAD_SINK[p.data_ptr()] = (v, weakref.ref(p))
        self.attached = True
        if self.world > 1 and self.flat.is_cuda:
            ops.GRAD_READY_HOOK = self._on_grad_ready
            # different ranks must not draw the same dropout masks for their shards (the seeds follow torch.manual_seed,
            # usually identical on every rank): mix the rank into the device salt once
            rank = dist.get_rank(self.group)
            ops.dropout_salt_set((ops.dropout_salt_get() + 0x9E3779B9 * (rank + 1)) & 0xFFFFFFFF)
        return self

    def detach(self):
        from . import ops
        for p in self.params:
            ops.GRAD_SINK.pop(p.data_ptr(), None)
        if ops.GRAD_READY_HOOK == self._on_grad_ready:
            ops.GRAD_READY_HOOK = None
        self.attached = False

    def __del__(self):
        try:
            if self.attached:
                self.detach()
        except Exception:
            pass

    def zero(self):
        self.flat.zero_()
        self._pending = []

    def set_weight(self, weight: float):
        """Each rank's loss is a mean over its LOCAL claims; weight = local_claims * world / global_claims turns the average
        over ranks into the gradient of the mean over the global batch (unequal shards)."""
        self._weight = float(weight)

    def _on_grad_ready(self, tag: str):
        """Called from the backward pass (ops.grad_marker) when every gradient of a chunk has been written: all-reduce that
        chunk on the communication stream while the backward pass continues on the compute stream."""
        c = self._chunk_index.get(CHUNK_TAGS.get(tag, -1))
        if (c is None or self.world <= 1 or not self.attached or not self.overlap or c in self._pending
                or c == len(self.chunks) - 1):
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.flat.device)
        cur = torch.cuda.current_stream()
        self._comm_stream.wait_stream(cur)
        lo, hi = self.chunks[c]
        with torch.cuda.stream(self._comm_stream):
            self._all_reduce(self.flat[lo:hi])
        self._pending.append(c)

    def reduce(self, weight: Optional[float] = None, collective: bool = True):
        """Average gradients over ranks. Leaves p.grad pointing into the bucket. collective=False only gathers the
        gradients into the bucket (warm-up steps that must not talk to other ranks)."""
        if weight is not None:
            self._weight = float(weight)
        if not self.attached:
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
            if self._pending:
                for c, (lo, hi) in enumerate(self.chunks):
                    if c not in self._pending:
                        self._all_reduce(self.flat[lo:hi])
                torch.cuda.current_stream().wait_stream(self._comm_stream)
            else:
                self._all_reduce(self.flat)
        elif self._pending and self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._pending = []
        if not self.attached:
            for p, v in zip(self.params, self.views):
                p.grad = v


class FlatAdam(object):
    """Adam with L2 weight decay (the reference fitter's torch.optim.Adam(lr, weight_decay=1e-3),
    Fitting/FittingFC/declare_fitter.py:57-61) as ONE kernel over flat buffers: the parameters are re-pointed at views of
    a flat fp32 buffer laid out like the reducer's gradient bucket, so param / grad / m / v are four flat arrays.
    Capturable: the step counter lives on the device."""

    def __init__(self, reducer: FlatGradAllReduce, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.reducer = reducer
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        flat = torch.zeros_like(reducer.flat)
        with torch.no_grad():
            for p, off in zip(reducer.params, reducer.offsets):
                v = flat[off:off + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
        self.flat_param = flat
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = torch.zeros((1,), dtype=torch.float32, device=flat.device)
        self.state = {}                  # (torch.optim.Optimizer look-alike for CapturedTrainStep's bookkeeping)
        if reducer.attached:
            reducer.attach()             # data pointers moved: re-register the gradient sinks

    def zero_grad(self, set_to_none: bool = False):
        self.reducer.zero()

    def step(self):
        from . import _lib, ops
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.get_adam_flat_f32(self.flat_param.data_ptr(), self.reducer.flat.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), self.flat_param.numel(), self.lr, self.betas[0], self.betas[1],
                                         self.eps, self.weight_decay, self.step_count.data_ptr(), st), "get_adam_flat_f32")
        ops.weights_updated()

    def state_tensors(self):
        return [self.flat_param, self.exp_avg, self.exp_avg_sq, self.step_count]
