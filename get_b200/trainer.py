"""Thin data-parallel trainer with the reference fitter's call convention.

Mirrors `CharManFitterQueryRepr1.fit` (Fitting/FittingFC/char_man_fitter_query_repr1.py:44-157) and `DeclareFitter._initialize`
(Fitting/FittingFC/declare_fitter.py:44-70): Adam(lr, weight_decay = reg_l2 = 1e-3) on the cross-entropy of the claim
logits, one optimizer step per mini-batch of claims, validation after every epoch, the best-validation `state_dict`
(reference key names, so the reference's `load_best_model` reads it) saved whenever macro-F1 improves, early stopping
by patience, `ValueError('Degenerate epoch loss')` on NaN / zero epoch loss.

What is different from the reference loop is only HOW a step runs: every mini-batch is sharded by claims over the ranks
of `torch.distributed` (balanced by evidence count, get_b200.ddp.shard_claims), each rank replays a captured CUDA
graph of [forward, loss, backward, overlapped gradient all-reduce, Adam] (get_b200.step_graph.CapturedTrainStep), and
validation runs batched (get_b200.evaluate). Launch with `python -m torch.distributed.run --nproc-per-node N ...` for N GPUs; a single
process trains on one GPU with the same code.

Mini-batches are numpy dicts in the fitter's flattened layout (get_b200.synthetic.make_batch documents the keys)."""
import os
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .ddp import FlatAdam, FlatGradAllReduce, balance_claims, trainable_named_parameters
from .keywords import KeyWordSettings as K
from .step_graph import CapturedTrainStep, pad_batch, select_claims


def classification_metrics(labels: Sequence[int], preds: Sequence[int], probs: Sequence[float]) -> Dict[str, float]:
    """The metrics of `_computing_metrics` (char_man_fitter_query_repr1.py:366-420) that drive model selection and reports:
    AUC of the class-1 logit, macro / micro F1, per-class precision / recall / F1."""
    y, p, s = np.asarray(labels), np.asarray(preds), np.asarray(probs, dtype=np.float64)
    out = {}
    pos, neg = s[y == 1], s[y == 0]
    if len(pos) and len(neg):          # AUC = P(score_pos > score_neg) + 0.5 P(tie)  (== sklearn roc_curve + auc)
        diff = pos[:, None] - neg[None, :]
        out["auc"] = float(((diff > 0).sum() + 0.5 * (diff == 0).sum()) / (len(pos) * len(neg)))
    else:
        out["auc"] = float("nan")
    f1s = []
    for c in (0, 1):
        tp = float(((p == c) & (y == c)).sum())
        fp = float(((p == c) & (y != c)).sum())
        fn = float(((p != c) & (y == c)).sum())
        prec = tp / (tp + fp) if tp + fp > 0 else 0.0
        rec = tp / (tp + fn) if tp + fn > 0 else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec > 0 else 0.0
        name = "true" if c == 1 else "false"
        out["precision_%s_cls" % name], out["recall_%s_cls" % name], out["f1_%s_cls" % name] = prec, rec, f1
        f1s.append(f1)
    out["f1_macro"] = float(np.mean(f1s))
    out["f1_micro"] = float((p == y).mean()) if len(y) else 0.0
    out["f1"] = out["f1_true_cls"]
    return out


class GETTrainer(object):
    def __init__(self, model, lr: float = 1e-4, reg_l2: float = 1e-3, n_iter: int = 100, early_stopping_patience: int = 10,
                 saved_model: Optional[str] = None, pad_pairs_to: int = 64, flat_adam: bool = True, log: Callable = None,
                 max_graphs: int = 32):
        self.model = model
        self.device = next(model.parameters()).device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        named = trainable_named_parameters(model)
        self.reducer = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
        if flat_adam:
            self.optimizer = FlatAdam(self.reducer, lr=lr, weight_decay=reg_l2)            # declare_fitter.py:57-61
        else:
            self.optimizer = torch.optim.Adam([p for _, p in named], lr=lr, weight_decay=reg_l2, fused=True, capturable=True)
        # batches are padded to multiples of `pad_pairs_to` pairs, so real Snopes batches (32..960 pairs) need at most 15 shapes
        self.stepper = CapturedTrainStep(model, self.optimizer, self.reducer, max_graphs=max_graphs)
        self.n_iter, self.patience, self.saved_model = int(n_iter), early_stopping_patience, saved_model
        self.pad_pairs_to = int(pad_pairs_to)
        self.log = log or (lambda *a: None)
        self.history: List[Dict] = []

    # ---------------------------------------------------------------------------------------------------------
    def train_step(self, batch: Dict) -> torch.Tensor:
        """One optimizer step on a GLOBAL mini-batch: this rank's claim shard goes through the captured step; returns the
        local loss (device scalar)."""
        from . import synthetic
        B = int(batch["query"].shape[0])
        mine = balance_claims(batch[K.EvidenceCountPerQuery], self.world)[self.rank] if self.world > 1 else list(range(B))
        local = pad_batch(select_claims(batch, mine), self.pad_pairs_to)
        q, d, l, kw = synthetic.batch_to_torch(local, device="cpu", pin=True)
        return self.stepper.step(q, d, l, kw, local["n_real_claims"], global_claims=B if self.world > 1 else 0)

    def fit(self, train_batches: Callable[[int], Iterable[Dict]], val_batches: Optional[Sequence[Dict]] = None) -> Dict:
        """train_batches(epoch) yields the (shuffled) global mini-batches of an epoch -- the same batches on every rank."""
        best_f1, best_epoch, bad = 0.0, 0, 0
        for epoch in range(self.n_iter):
            self.model.train(True)
            losses = [self.train_step(b) for b in train_batches(epoch)]
            epoch_loss = float(torch.stack(losses).sum().item()) if losses else 0.0     # one host sync per epoch
            rec = {"epoch": epoch, "epoch_loss": epoch_loss, "steps": len(losses)}
            if val_batches is not None:
                rec.update({"val_" + k: v for k, v in self.evaluate(val_batches).items()})
                f1 = rec["val_f1_macro"]
                if f1 > best_f1:
                    best_f1, best_epoch, bad = f1, epoch, 0
                    if self.saved_model and self.rank == 0:      # char_man_fitter_query_repr1.py:140-146
                        os.makedirs(os.path.dirname(os.path.abspath(self.saved_model)), exist_ok=True)
                        with open(self.saved_model, "wb") as fh:
                            torch.save(self.model.state_dict(), fh)
                else:
                    bad += 1
            self.history.append(rec)
            self.log(rec)
            if np.isnan(epoch_loss) or epoch_loss == 0.0:
                raise ValueError("Degenerate epoch loss: {}".format(epoch_loss))
            if val_batches is not None and self.patience and bad > self.patience:
                self.log("Early Stopped due to no better performance in %s epochs" % bad)
                break
        return {"best_val_f1_macro": best_f1, "best_epoch": best_epoch, "history": self.history}

    @torch.no_grad()
    def evaluate(self, batches: Sequence[Dict]) -> Dict[str, float]:
        """Batched evaluation (the reference goes claim by claim, char_man_fitter_query_repr1.py:278-360): prediction =
        argmax of the logits, "probability" = the raw class-1 logit."""
        from . import synthetic
        from .evaluate import predict_batched
        was_training = self.model.training
        labels, preds, probs = [], [], []
        for b in batches:
            q, d, l, kw = synthetic.batch_to_torch(b, device=self.device)
            pr, p1, _ = predict_batched(self.model, q, d, kw)
            labels += b["labels"].tolist()
            preds += pr.cpu().tolist()
            probs += p1.cpu().tolist()
        self.model.train(was_training)
        return classification_metrics(labels, preds, probs)

    def load_best_model(self):
        """char_man_fitter_query_repr1.py:524-530"""
        self.model.load_state_dict(torch.load(self.saved_model, map_location=self.device))
        ops.weights_updated()
