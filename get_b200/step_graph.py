"""Training step of the GET hot path captured into CUDA graphs (one graph per mini-batch shape).

At the reference's batch size (32 claims, ~200 claim-evidence pairs) one step is ~110 kernel launches of a few tens of
microseconds each: issued one by one from Python the step is bound by the host, not by the GPU. `CapturedTrainStep`
records  [advance dropout salt -> forward -> cross-entropy -> backward -> (gradient all-reduce) -> (optimizer step)]
once per shape and replays it with one launch; inputs are copied into static device buffers first.

The fitter hands over B1 = sum of evidence counts flattened pairs, which varies from batch to batch. To bound the number
of graphs the batch is padded with dummy claims (copies of the first claim with copies of its first evidence) up to a
multiple of `pad_pairs_to` pairs; dummy claims are excluded from the loss, so the real claims' logits, the loss and all
gradients are unchanged (every claim is independent of the others).

Dropout: a replay repeats the captured kernel arguments, so the per-step variation of the masks comes from the
library's device salt word, advanced by the first node of the graph (include/get_b200.h, get_dropout_salt_advance).
"""
import collections
import os
import copy
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops
from .keywords import KeyWordSettings as K

_TENSOR_KEYS = (K.Query_lens, K.QuerySources, K.DocSources, K.DocContentNoPaddingEvidence, K.EvidenceCountPerQuery,
                K.Query_Adj, K.Evd_Docs_Adj)


def pad_batch(batch: Dict, multiple: int) -> Dict:
    """numpy mini-batch (get_b200.synthetic.make_batch layout / the fitter's flattened layout) -> the same batch with
    dummy claims appended so that the number of pairs is a multiple of `multiple`. Adds 'n_real_claims', 'real_pairs'."""
    B1 = int(batch[K.DocContentNoPaddingEvidence].shape[0])
    B = int(batch["query"].shape[0])
    n = int(batch[K.FIXED_NUM_EVIDENCES])
    out = dict(batch)
    out["n_real_claims"], out["real_pairs"] = B, B1
    pad = (-B1) % max(1, int(multiple))
    if pad == 0:
        return out
    counts = []
    while pad > 0:
        c = min(n, pad)
        counts.append(c)
        pad -= c
    nd, extra = len(counts), int(sum(counts))

    def rep0(a, reps):               # `reps` copies of the first row
        return np.repeat(a[:1], reps, axis=0)

    out["query"] = np.concatenate([batch["query"], rep0(batch["query"], nd)])
    out["labels"] = np.concatenate([batch["labels"], np.zeros((nd,), batch["labels"].dtype)])
    out[K.Query_lens] = np.concatenate([batch[K.Query_lens], rep0(batch[K.Query_lens], nd)])
    out[K.Query_Adj] = np.concatenate([batch[K.Query_Adj], rep0(batch[K.Query_Adj], nd)])
    out[K.QuerySources] = np.concatenate([batch[K.QuerySources], rep0(batch[K.QuerySources], nd)])
    doc = np.zeros((nd,) + batch["document"].shape[1:], batch["document"].dtype)
    src = np.full((nd, n), -1, batch[K.DocSources].dtype)
    lens = np.zeros((nd, n), batch[K.Doc_lens].dtype)
    for i, c in enumerate(counts):
        doc[i, :c] = batch["document"][0, 0]
        src[i, :c] = batch[K.DocSources][0, 0]
        lens[i, :c] = batch[K.Doc_lens][0, 0]
    out["document"] = np.concatenate([batch["document"], doc])
    out[K.DocSources] = np.concatenate([batch[K.DocSources], src])
    out[K.Doc_lens] = np.concatenate([batch[K.Doc_lens], lens])
    out[K.EvidenceCountPerQuery] = np.concatenate([batch[K.EvidenceCountPerQuery],
                                                   np.asarray(counts, batch[K.EvidenceCountPerQuery].dtype)])
    out[K.DocContentNoPaddingEvidence] = np.concatenate([batch[K.DocContentNoPaddingEvidence],
                                                         rep0(batch[K.DocContentNoPaddingEvidence], extra)])
    out[K.Evd_Docs_Adj] = np.concatenate([batch[K.Evd_Docs_Adj], rep0(batch[K.Evd_Docs_Adj], extra)])
    out["e_lens"] = np.concatenate([batch["e_lens"], rep0(batch["e_lens"], extra)])
    if "raw_doc_tokens" in batch:
        out["raw_query_tokens"] = np.concatenate([batch["raw_query_tokens"], rep0(batch["raw_query_tokens"], nd)])
        out["raw_query_lens"] = np.concatenate([batch["raw_query_lens"], rep0(batch["raw_query_lens"], nd)])
        out["raw_doc_tokens"] = np.concatenate([batch["raw_doc_tokens"], rep0(batch["raw_doc_tokens"], extra)])
        out["raw_doc_lens"] = np.concatenate([batch["raw_doc_lens"], rep0(batch["raw_doc_lens"], extra)])
    out["pairs"] = B1 + extra
    return out


def slice_batch(batch: Dict, lo: int, hi: int) -> Dict:
    """Claims [lo, hi) of a numpy mini-batch in the fitter's flattened layout (this rank's shard of a global batch): per-claim
    arrays are sliced by claim, the flattened evidence arrays by the prefix sums of the evidence counts."""
    cnt = np.asarray(batch[K.EvidenceCountPerQuery]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(cnt)])
    r0, r1 = int(off[lo]), int(off[hi])
    out = dict(batch)
    for k in ("query", "document", "labels", K.Query_lens, K.Query_Adj, K.QuerySources, K.DocSources, K.Doc_lens,
              K.EvidenceCountPerQuery, "raw_query_tokens", "raw_query_lens"):
        if k in batch:
            out[k] = batch[k][lo:hi]
    for k in (K.DocContentNoPaddingEvidence, K.Evd_Docs_Adj, "e_lens", "raw_doc_tokens", "raw_doc_lens"):
        if k in batch:
            out[k] = batch[k][r0:r1]
    out["pairs"] = r1 - r0
    return out


def select_claims(batch: Dict, claims) -> Dict:
    """The claims `claims` (any subset, any order) of a numpy mini-batch in the fitter's flattened layout: per-claim arrays
    are gathered by claim, the flattened evidence arrays by the evidence rows of those claims (a mini-batch is a SET of
    claims: the mean loss and its gradient do not depend on the order)."""
    claims = np.asarray(claims, dtype=np.int64)
    cnt = np.asarray(batch[K.EvidenceCountPerQuery]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(cnt)])
    rows = np.concatenate([np.arange(off[c], off[c + 1]) for c in claims]) if len(claims) else np.zeros((0,), np.int64)
    out = dict(batch)
    for k in ("query", "document", "labels", K.Query_lens, K.Query_Adj, K.QuerySources, K.DocSources, K.Doc_lens,
              K.EvidenceCountPerQuery, "raw_query_tokens", "raw_query_lens"):
        if k in batch:
            out[k] = batch[k][claims]
    for k in (K.DocContentNoPaddingEvidence, K.Evd_Docs_Adj, "e_lens", "raw_doc_tokens", "raw_doc_lens"):
        if k in batch:
            out[k] = batch[k][rows]
    out["pairs"] = int(len(rows))
    return out


def token_batch_to_host(batch: Dict, pin: bool = True):
    """The compact host-side form of a mini-batch: raw token ids + counts + sources + labels, a few
    hundred kB instead of the 18.8 MB of dense float64 adjacencies. Returns a dict of (pinned) CPU tensors."""
    def t(x):
        y = torch.from_numpy(np.ascontiguousarray(x))
        return y.pin_memory() if pin and torch.cuda.is_available() else y
    return {"q_tok": t(batch["raw_query_tokens"]), "q_len": t(batch["raw_query_lens"]), "d_tok": t(batch["raw_doc_tokens"]),
            "d_len": t(batch["raw_doc_lens"]), "cnt": t(batch[K.EvidenceCountPerQuery]), "labels": t(batch["labels"]),
            "q_src": t(batch[K.QuerySources]), "d_src": t(batch[K.DocSources]),
            "n": int(batch[K.FIXED_NUM_EVIDENCES]), "window": int(batch["window"]),
            "L": int(batch["query"].shape[1]), "R": int(batch[K.DocContentNoPaddingEvidence].shape[1])}


def device_batch_from_tokens(tb: Dict, device):
    """Compact host batch -> (query, document, labels, kwargs) on the device in the fitter's calling convention, with the
    word graphs built by get_build_word_graphs (no host syncs: every shape comes from the host-side tensors)."""
    from .graph_build import build_word_graphs
    mv = lambda x: x.to(device, non_blocking=True)
    q_nodes, q_adj, q_n = build_word_graphs(mv(tb["q_tok"]), mv(tb["q_len"]), tb["L"], tb["window"])
    d_nodes, d_adj, d_n = build_word_graphs(mv(tb["d_tok"]), mv(tb["d_len"]), tb["R"], tb["window"])
    cnt = mv(tb["cnt"])
    B, n, R, b1 = tb["q_tok"].shape[0], tb["n"], tb["R"], tb["d_tok"].shape[0]
    seg = torch.repeat_interleave(torch.arange(B, device=device), cnt, output_size=b1)
    off = torch.cumsum(cnt, 0) - cnt
    slot = seg * n + (torch.arange(b1, device=device) - off[seg])
    document = torch.zeros((B * n, R), dtype=torch.int64, device=device)
    document.index_copy_(0, slot, d_nodes)
    kw = {K.Query_lens: q_n.to(torch.int64), K.Doc_lens: None, K.DocLensIndices: (None, None, d_n.to(torch.int64)),
          K.QueryLensIndices: (None, None, q_n.to(torch.int64)), K.QuerySources: mv(tb["q_src"]), K.DocSources: mv(tb["d_src"]),
          K.DocContentNoPaddingEvidence: d_nodes, K.EvidenceCountPerQuery: cnt, K.FIXED_NUM_EVIDENCES: n,
          K.Query_Adj: q_adj, K.Evd_Docs_Adj: d_adj}
    return q_nodes, document.view(B, n, R), mv(tb["labels"]), kw


class _Slot(object):
    __slots__ = ("graph", "query", "document", "labels", "kw", "loss", "logits", "n_real", "launches", "global_claims")


class CapturedTrainStep(object):
    """step(query, document, labels, kwargs, n_real_claims) -> loss (0-d device tensor, valid until the next step).

    All tensor arguments may live on the host (pinned for asynchronous copies) or on the device; shapes select the graph.
    `optimizer` (capturable) and `reducer` (get_b200.ddp.FlatGradAllReduce) are optional parts of the captured step.
    Multi-GPU: every call of step() executes exactly one gradient all-reduce (inside the replayed graph), whether or not
    the shape is new on this rank, so ranks stay in lock-step as long as they call step() equally often."""

    def __init__(self, model, optimizer=None, reducer=None, loss_fn=None, collective_in_graph: bool = True,
                 accumulate: bool = False, max_graphs: int = 32):
        """accumulate=True (micro-batching, needs an attached reducer): the captured step is [forward, loss * loss_scale,
        backward] only -- gradients accumulate in the reducer's bucket across calls; the caller zeroes the bucket, runs the
        all-reduce and the optimizer once per optimizer step."""
        self.model, self.optimizer, self.reducer = model, optimizer, reducer
        self.accumulate = bool(accumulate)
        self.loss_scale = 1.0
        assert not accumulate or (reducer is not None and optimizer is None)
        # collective_in_graph=False: the graph ends after the backward pass (gradients gathered in the flat bucket); the
        # all-reduce and the optimizer step are then issued eagerly after every replay
        if os.environ.get("GET_B200_SPLIT_TAIL", "0") == "1":
            collective_in_graph = False
        self.split_tail = (not collective_in_graph) and reducer is not None and reducer.world > 1
        self.loss_fn = loss_fn or ops.cross_entropy
        # one graph (with its private pool of activations, ~4 MB per claim-evidence pair) per padded batch shape: the cache is
        # bounded, least-recently-used shapes are dropped and re-captured on demand (real Snopes batches span 32..960 pairs)
        self.slots: "collections.OrderedDict[Tuple, _Slot]" = collections.OrderedDict()
        self.max_graphs = max(1, int(max_graphs))
        self.evictions = 0
        self.device = next(model.parameters()).device
        if reducer is not None:
            reducer.init_collective()    # communicator created here, never inside a capture
            if not reducer.attached:
                reducer.attach()         # the flat bucket becomes the gradient storage (weight-gradient kernels write into it)
        self.replayed_launches = 0       # kernels of libget_b200.so launched through graph replays so far

    # ---------------------------------------------------------------------------------------------------------
    def _key(self, query, document, kw, n_real, global_claims=0):
        return (tuple(query.shape), tuple(document.shape), tuple(kw[K.DocContentNoPaddingEvidence].shape),
                tuple(kw[K.Evd_Docs_Adj].shape), str(kw[K.Evd_Docs_Adj].dtype), int(n_real), bool(self.model.training),
                int(global_claims))

    def _eager(self, s: _Slot, collective: bool = True):
        red = self.reducer
        in_step = collective and not self.split_tail and not self.accumulate
        if red is not None:
            # chunk all-reduces are issued from the backward pass (GET_B200_NO_OVERLAP=1: one all-reduce after it, for A/B runs)
            red.overlap = in_step and os.environ.get("GET_B200_NO_OVERLAP", "0") != "1"
            # the local loss is a mean over the LOCAL claims: re-weight unequal shards
            red.set_weight(s.n_real * red.world / s.global_claims if (s.global_claims and red.world > 1) else 1.0)
        logits = self.model(s.query, s.document, **s.kw)
        loss = self.loss_fn(logits[:s.n_real], s.labels[:s.n_real])
        (loss if self.loss_scale == 1.0 else loss * self.loss_scale).backward()
        # detached: a slot must not keep the autograd graph alive across steps -- a live graph pins the parameters'
        # AccumulateGrad nodes to the stream they were created on (an earlier capture / warm-up stream), and the next
        # capture then forks into that stale stream
        if self.accumulate:
            return logits.detach(), loss.detach()
        if red is not None:
            red.reduce(collective=in_step)
        if self.optimizer is not None and not self.split_tail:
            self.optimizer.step()
        return logits.detach(), loss.detach()

    def _tail(self):
        """all-reduce + optimizer step outside the graph (collective_in_graph=False)."""
        self.reducer.reduce(collective=True)
        if self.optimizer is not None:
            self.optimizer.step()

    def _build(self, query, document, labels, kw, n_real, global_claims=0) -> _Slot:
        dev = self.device
        s = _Slot()
        s.n_real = int(n_real)
        s.global_claims = int(global_claims)
        new = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)
        s.query, s.document, s.labels = new(query), new(document), new(labels)
        s.kw = dict(kw)
        for k in _TENSOR_KEYS:
            if k in kw and torch.is_tensor(kw[k]):
                s.kw[k] = new(kw[k])
        e_lens = kw[K.DocLensIndices][2]
        s.kw[K.DocLensIndices] = (None, None, new(e_lens))
        if K.QueryLensIndices in kw:
            s.kw[K.QueryLensIndices] = (None, None, s.kw[K.Query_lens])
        self._copy_in(s, query, document, labels, kw)
        # ---- one eager warm-up step on a side stream (lazy initialisation must not happen inside the capture);
        #      parameters, optimizer state and the dropout salt are restored afterwards: building a graph is not a step
        params = [p for p in self.model.parameters()]
        backup = [p.detach().clone() for p in params]
        # optimizer state is restored IN PLACE: graphs captured earlier hold the addresses of these tensors
        opt_backup = flat_backup = None
        if self.optimizer is not None:
            opt_backup = {p: {k: (v.detach().clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in st.items()}
                          for p, st in self.optimizer.state.items()}
            if hasattr(self.optimizer, "state_tensors"):        # get_b200.ddp.FlatAdam: flat moment buffers + device step
                flat_backup = [t.detach().clone() for t in self.optimizer.state_tensors()]
        salt = ops.dropout_salt_get()
        bucket_backup = self.reducer.flat.clone() if self.accumulate else None
        cur = torch.cuda.current_stream()
        if getattr(self, "_warm_stream", None) is None:
            self._warm_stream = torch.cuda.Stream()       # ONE warm-up stream for all builds (streams come from a small pool)
        side = self._warm_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if not self.accumulate:
                self._zero_grad()
            self._eager(s, collective=False)    # ranks see different shapes: building a graph must not be a collective
        cur.wait_stream(side)
        torch.cuda.synchronize()
        if bucket_backup is not None:
            self.reducer.flat.copy_(bucket_backup)     # building a graph is not a step: gradients accumulated so far stay
        with torch.no_grad():
            for p, b in zip(params, backup):
                p.copy_(b)
        if self.optimizer is not None:
            with torch.no_grad():
                for p, st in self.optimizer.state.items():
                    old = opt_backup.get(p)
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()          # state created by the warm-up step: back to its initial value
                        elif old is not None and k in old:
                            st[k] = old[k]
            if flat_backup is not None:
                with torch.no_grad():
                    for t, b in zip(self.optimizer.state_tensors(), flat_backup):
                        t.copy_(b)
                ops.weights_updated()
        ops.dropout_salt_set(salt)
        # ---- capture
        ops.prepare_split_table()      # host -> device copy of the weight-split job table: not allowed inside the capture
        if not self.accumulate:
            self._zero_grad()
        prof, ops.PROFILE_GSL_EVENTS = ops.PROFILE_GSL_EVENTS, None
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # other threads of a multi-GPU process (the NCCL watchdog polls its events) must not invalidate the capture
        multi = self.reducer is not None and self.reducer.world > 1
        if multi:
            torch.cuda.synchronize()
        mode = os.environ.get("GET_B200_CAPTURE_MODE", "thread_local" if multi else "global")
        with torch.cuda.graph(g, capture_error_mode=mode):
            ops.begin_step_capture()
            ops.dropout_salt_advance()
            if self.reducer is not None and self.reducer.attached and not self.accumulate:
                self.reducer.zero()        # part of the replayed step: gradients accumulate into the bucket
            s.logits, s.loss = self._eager(s)
        s.launches = _lib.launch_count() - n0      # kernels of libget_b200.so recorded in this graph
        ops.PROFILE_GSL_EVENTS = prof
        s.graph = g
        return s

    def _zero_grad(self):
        if self.reducer is not None and self.reducer.attached:
            self.reducer.zero()            # p.grad are views of the flat bucket: one memset, the views stay in place
        elif self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            self.model.zero_grad(set_to_none=True)

    @staticmethod
    def _copy_in(s: _Slot, query, document, labels, kw):
        s.query.copy_(query, non_blocking=True)
        s.document.copy_(document, non_blocking=True)
        s.labels.copy_(labels, non_blocking=True)
        for k in _TENSOR_KEYS:
            if k in kw and torch.is_tensor(kw[k]):
                s.kw[k].copy_(kw[k], non_blocking=True)
        s.kw[K.DocLensIndices][2].copy_(kw[K.DocLensIndices][2], non_blocking=True)

    # ---------------------------------------------------------------------------------------------------------
    def step(self, query, document, labels, kw, n_real_claims: Optional[int] = None, global_claims: int = 0) -> torch.Tensor:
        """global_claims (multi-GPU, unequal shards): real claims of the GLOBAL batch; the local gradients are re-weighted by
        n_real * world / global_claims so that the all-reduced average is the gradient of the global-batch mean loss."""
        n_real = int(n_real_claims if n_real_claims is not None else query.shape[0])
        key = self._key(query, document, kw, n_real, global_claims)
        s = self.slots.get(key)
        if s is None:
            while len(self.slots) >= self.max_graphs:
                torch.cuda.synchronize()            # nothing in flight may still replay the graph that is dropped
                _, old = self.slots.popitem(last=False)
                del old
                self.evictions += 1
            s = self._build(query, document, labels, kw, n_real, global_claims)
            self.slots[key] = s
            # the capture itself executes nothing: fall through and replay once so that this call IS a step
        else:
            self.slots.move_to_end(key)
            self._copy_in(s, query, document, labels, kw)
        s.graph.replay()
        self.replayed_launches += s.launches
        if self.split_tail:
            self._tail()
        elif self.optimizer is not None:
            # the replay changed the weights on the device without any host-side trace (no version counter, no optimizer
            # hook): mark every packed weight stale so that an EAGER forward issued next (evaluation) re-packs them
            ops.weights_updated()
        return s.loss

    # ---------------------------------------------------------------------------------------------------------
    def prefetch(self, query, document, labels, kw):
        """Start copying the NEXT batch (host, pinned) to device staging buffers on a side stream while the current step
        runs; returns a handle to pass to step_prefetched(). The staging -> static copy is device to device (microseconds)."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = {}
        key = (tuple(query.shape), tuple(document.shape), tuple(kw[K.DocContentNoPaddingEvidence].shape),
               str(kw[K.Evd_Docs_Adj].dtype))
        # two staging sets per shape, used alternately: the set filled now is not the one the running step reads from
        sets = self._staging.setdefault(key, [None, None, 0])
        slot = sets[2] & 1
        sets[2] += 1
        dev = self.device
        new = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)
        if sets[slot] is None:
            st = {"query": new(query), "document": new(document), "labels": new(labels),
                  "e_lens": new(kw[K.DocLensIndices][2]), "event": torch.cuda.Event()}
            for k in _TENSOR_KEYS:
                if k in kw and torch.is_tensor(kw[k]):
                    st[k] = new(kw[k])
            sets[slot] = st
        st = sets[slot]
        if st.get("consumed") is not None:         # the step that last read this staging set has copied it out
            self._copy_stream.wait_event(st["consumed"])
        with torch.cuda.stream(self._copy_stream):
            st["query"].copy_(query, non_blocking=True)
            st["document"].copy_(document, non_blocking=True)
            st["labels"].copy_(labels, non_blocking=True)
            st["e_lens"].copy_(kw[K.DocLensIndices][2], non_blocking=True)
            for k in _TENSOR_KEYS:
                if k in kw and torch.is_tensor(kw[k]):
                    st[k].copy_(kw[k], non_blocking=True)
            st["event"].record(self._copy_stream)
        kw_dev = dict(kw)
        for k in _TENSOR_KEYS:
            if k in st:
                kw_dev[k] = st[k]
        kw_dev[K.DocLensIndices] = (None, None, st["e_lens"])
        return (st, kw_dev)

    def prefetch_tokens(self, tb: Dict):
        """Compact host batch (token_batch_to_host) -> device batch on the copy stream while the current step runs (SURVEY.md
        8f rank 1): the H2D copy of the token ids (~0.2 MB) and the construction of node lists and normalised adjacencies
        (get_build_word_graphs) overlap the running step; returns a handle for step_prefetched()."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = {}
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            q, d, l, kw = device_batch_from_tokens(tb, self.device)
            ev = torch.cuda.Event()
            ev.record()
        e_lens = kw[K.DocLensIndices][2]
        for t in [q, d, l, e_lens] + [v for v in kw.values() if torch.is_tensor(v)]:
            t.record_stream(main)      # allocated on the copy stream, read by the main stream's copy into the static inputs
        return ({"query": q, "document": d, "labels": l, "e_lens": e_lens, "event": ev}, kw)

    def step_prefetched(self, handle, n_real_claims: Optional[int] = None, global_claims: int = 0) -> torch.Tensor:
        """Call order for full overlap: loss = step_prefetched(h_i); h_next = prefetch(batch_{i+1}); read loss."""
        st, kw_dev = handle
        torch.cuda.current_stream().wait_event(st["event"])
        loss = self.step(st["query"], st["document"], st["labels"], kw_dev, n_real_claims, global_claims)
        if st.get("consumed") is None:
            st["consumed"] = torch.cuda.Event()
        st["consumed"].record()       # (recorded after the replay: a little late, but it never delays the next-but-one copy)
        return loss

    def n_graphs(self) -> int:
        return len(self.slots)
