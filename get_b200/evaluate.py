"""Batched inference for the GET hot path.

The reference evaluates ONE claim per forward (`CharManFitterQueryRepr1.evaluate`,
Fitting/FittingFC/char_man_fitter_query_repr1.py:278-360: B=1, B1=n_c, `net.predict`, prediction = argmax of the logits,
"probability" for AUC = the raw class-1 logit :353-360). Every (claim, evidence set) is independent, so the same numbers
come out of one forward over many claims; `CapturedForward` additionally replays that forward as a CUDA graph per shape.
"""
from typing import Dict, Optional, Tuple

import torch

from .keywords import KeyWordSettings as K
from .step_graph import _TENSOR_KEYS


class CapturedForward(object):
    """forward(query, document, kwargs) -> logits (static device tensor, valid until the next call). Eval mode only."""

    def __init__(self, model):
        self.model = model
        self.device = next(model.parameters()).device
        self.slots: Dict[Tuple, dict] = {}

    def _copy_in(self, s, query, document, kw):
        s["query"].copy_(query, non_blocking=True)
        s["document"].copy_(document, non_blocking=True)
        for k in _TENSOR_KEYS:
            if k in kw and torch.is_tensor(kw[k]):
                s["kw"][k].copy_(kw[k], non_blocking=True)
        s["kw"][K.DocLensIndices][2].copy_(kw[K.DocLensIndices][2], non_blocking=True)

    def forward(self, query, document, kw) -> torch.Tensor:
        assert not self.model.training, "CapturedForward is an inference path: call model.eval() first"
        key = (tuple(query.shape), tuple(document.shape), tuple(kw[K.DocContentNoPaddingEvidence].shape),
               str(kw[K.Evd_Docs_Adj].dtype))
        s = self.slots.get(key)
        if s is None:
            dev = self.device
            new = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)
            s = {"query": new(query), "document": new(document), "kw": dict(kw)}
            for k in _TENSOR_KEYS:
                if k in kw and torch.is_tensor(kw[k]):
                    s["kw"][k] = new(kw[k])
            s["kw"][K.DocLensIndices] = (None, None, new(kw[K.DocLensIndices][2]))
            s["kw"].pop(K.OutputRankingKey, None)
            self._copy_in(s, query, document, kw)
            cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                self.model(s["query"], s["document"], **s["kw"])          # lazy initialisation outside the capture
            cur.wait_stream(side)
            torch.cuda.synchronize()
            from . import ops
            ops.prepare_split_table()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g), torch.no_grad():
                ops.begin_step_capture()       # the weights may have changed since the graph was captured: re-split inside
                s["logits"] = self.model(s["query"], s["document"], **s["kw"])
            s["graph"] = g
            self.slots[key] = s
        else:
            self._copy_in(s, query, document, kw)
        s["graph"].replay()
        return s["logits"]


@torch.no_grad()
def predict_batched(model, query, document, kw, captured: Optional[CapturedForward] = None):
    """Many claims in one forward. Returns (predicted class (B,), class-1 logit (B,), logits (B, C)) -- the quantities
    the reference's `evaluate` derives claim by claim (char_man_fitter_query_repr1.py:349-360)."""
    model.train(False)
    logits = captured.forward(query, document, kw) if captured is not None else model(query, document, **kw)
    return torch.argmax(logits, dim=1), logits[:, 1], logits
