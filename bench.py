#!/usr/bin/env python
"""Benchmark of the GET hot path (BASELINE.json metric: claim-evidence pairs/s, forward+backward).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU path (oracle port)

A "step" = one training pass over one mini-batch of the Snopes configuration (BASELINE.json configs[1]:
B=32 claims, L=30, R=100, D=H=300, heads 5/2, window 3, gsl_rate 0.6, fp32): forward, cross-entropy, backward,
gradient all-reduce (N>1) and the Adam update of the reference fitter (lr 1e-4, weight_decay 1e-3,
Fitting/FittingFC/declare_fitter.py:57-61). Per-GPU work is fixed (weak scaling): every rank processes its own
stream of 32-claim batches. Inputs rotate over NBATCH distinct pre-generated batches.

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NBATCH = 16
PAD_PAIRS = 16
WORKLOAD = "snopes"
METRIC = "claim_evidence_pairs_per_sec_fwd_bwd"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlClockSampler(object):
    """SM clock and throttle reasons read through NVML every ~2 ms DURING the timed region (the timed region of the default
    run lasts ~0.1 s, too short for nvidia-smi's polling). Falls back to the nvidia-smi sampler below when NVML is missing."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        try:                      # NVML ignores CUDA_VISIBLE_DEVICES: address the device by UUID when torch exposes it
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) and bytes is not str else uuid)
        except Exception:
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.sm, self.mask, self.stop_flag, self.thread = [], 0, False, None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "samples": len(self.sm),
                "reasons": reasons, "source": "nvml"}


def make_clock_sampler(index: int):
    try:
        return NvmlClockSampler(index)
    except Exception:
        return ClockSampler(index)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_REAL_STDOUT = None


def guard_stdout():
    """stdout must carry exactly ONE JSON line: point fd 1 at stderr for the whole run (library banners such as NCCL's
    version line, warnings, device printf) and restore it only to emit the result."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line))
    sys.stdout.flush()


def make_batches(w, n, base_seed, n_claims=None):
    from get_b200 import synthetic
    return [synthetic.make_batch(w, seed=base_seed + 1000 * i, n_claims=n_claims) for i in range(n)]


def graph_kernel_bytes(w, pairs):
    """SURVEY.md 8(d)'s per-pair figure for the GSL path: read F1 + the dense adjacency, write the refined aggregation
    (two bf16 planes = 4 bytes per element, like the fp32 the formula was written for): 4*R*(2H+R) per pair. Since round 2
    the adjacency is read by the list builder (once per step, for five consumers), not by the fused kernel."""
    return pairs * 4 * w.len_right * (2 * w.hidden + w.len_right)


def gsl_kernel_own_bytes(w, pairs, nnz_per_pair, n_sp, planes):
    """Bytes the fused GSL kernel itself has to move per launch: F1 read once (4RH), refined aggregation written once
    (`planes` bf16 planes, or fp32 when planes = 0), neighbour lists (8 B per edge + 4 B per row), scorer projection partial
    sums (4 B x n_sp per row), score + keep written (5 B per row)."""
    R, H = w.len_right, w.hidden
    out = 2 * planes * R * (((H + 7) // 8) * 8) if planes else 4 * R * H
    return pairs * (4 * R * H + out + 8 * nnz_per_pair + 4 * R + 4 * n_sp * R + 5 * R)


def config_dict(w):
    """Workload description shared VERBATIM by both arms (the driver compares it)."""
    return {"workload": "%s B=%d L=%d R=%d D=%d H=%d heads=%d/%d window=%d gsl_rate=%.1f" % (
        w.name, w.batch_claims, w.len_left, w.len_right, w.emb_dim, w.hidden, w.heads_words, w.heads_evds, w.window,
        w.gsl_rate),
        "step": "forward + cross-entropy + backward + grad all-reduce (N>1) + Adam(lr=1e-4, wd=1e-3), train-mode dropout",
        "claims_per_gpu_per_step": w.batch_claims,
        "l2": "inputs rotate over %d distinct batches; every step also writes ~0.8 GB of activations (> 126 MB L2)" % NBATCH}


# =================================================================================================
# the reference: its own unmodified modules (oracle/_ref, a verbatim copy made by oracle/make_ref.py) when available,
# else the oracle port
# =================================================================================================
class _cpu_shim(object):
    """Models/BiDAF/wrapper.py:221 hard-codes `.cuda()`: while the reference runs on the host cores `Tensor.cuda` is the
    identity (restored afterwards, so the same process can also run the reference on the GPU)."""

    def __enter__(self):
        import torch
        self.t, self.orig = torch, torch.Tensor.cuda
        torch.Tensor.cuda = lambda self_, *a, **k: self_

    def __exit__(self, *exc):
        self.t.Tensor.cuda = self.orig


def reference_modules():
    from oracle import ref_import        # bench.py's reference / baseline legs may execute oracle/
    if not ref_import.available():
        return None
    return ref_import.import_reference()[0]


def reference_step_fn(w, device, train=True):
    """(step(i) -> None, forward(i) -> None, pairs per batch list): the reference's own training step -- zero_grad, forward,
    losses.cross_entroy, backward, Adam(lr, weight_decay=1e-3) step (declare_fitter.py:57-61, char_man_fitter...:92-128)."""
    import torch
    from get_b200 import synthetic
    gbss = reference_modules()
    torch.manual_seed(123756)
    model = gbss.Graph_basedSemantiStructure(synthetic.match_params(w, cuda=(device != "cpu")))
    model = model.to(device)
    model.train(train)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3)
    batches = make_batches(w, 4, 123756)
    tens = [synthetic.batch_to_torch(b, device=device) for b in batches]
    lossf = torch.nn.CrossEntropyLoss()           # losses.py:29-32

    def step(i):
        q, d, l, kw = tens[i % len(tens)]
        opt.zero_grad()
        loss = lossf(model(q, d, **kw), l.long())
        loss.backward()
        opt.step()

    def forward(i):
        q, d, l, kw = tens[i % len(tens)]
        with torch.no_grad():
            model(q, d, **kw)
    return step, forward, [b["pairs"] for b in batches], model


def cpu_baseline(w, steps, warmup, cores, max_seconds=40.0):
    """The reference on the host cores: its own modules in train mode (kind "reference") or, when oracle/_ref is absent, the
    oracle port (eval-mode forward + backward; kind "port")."""
    import torch
    torch.set_num_threads(cores)
    if reference_modules() is not None:
        with _cpu_shim():
            step, _, pairs_list, _ = reference_step_fn(w, "cpu", train=True)
            pairs, t_total, done = 0, 0.0, 0
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                step(i)
                dt = time.perf_counter() - t0
                if i >= warmup:
                    t_total += dt
                    pairs += pairs_list[i % len(pairs_list)]
                    done += 1
                    if t_total > max_seconds:
                        break
        return {"value": pairs / t_total, "ms_per_step": 1e3 * t_total / done, "kind": "reference",
                "sample": "%d training steps (fwd + CE + bwd + Adam, train-mode dropout) of the UNMODIFIED reference modules on "
                          "%s batches (B=%d), %.1f s of CPU work" % (done, w.name, w.batch_claims, t_total)}
    from get_b200 import synthetic
    from get_b200.model import Graph_basedSemantiStructure
    from oracle import get_oracle as O
    torch.manual_seed(123756)
    model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=False))   # parameter container only (CPU)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = dict(gsl_rate=w.gsl_rate, use_claim_source=w.use_claim_source, use_article_source=w.use_article_source)
    batches = make_batches(w, min(4, max(1, steps)), 123756)
    tens = [synthetic.batch_to_torch(b) for b in batches]
    pairs, t_total, done = 0, 0.0, 0
    for i in range(warmup + steps):
        q, d, l, kw = tens[i % len(tens)]
        t0 = time.perf_counter()
        O.loss_and_grads(sd, cfg, q, d, l, kw)
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_total += dt
            pairs += batches[i % len(tens)]["pairs"]
            done += 1
            if t_total > max_seconds:
                break
    return {"value": pairs / t_total, "ms_per_step": 1e3 * t_total / done, "kind": "port",
            "sample": "%d fwd+bwd steps of the oracle port on %s batches (B=%d, eval-mode dropout), %.1f s of CPU work"
                      % (done, w.name, w.batch_claims, t_total)}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores, all threads."""
    import torch
    from get_b200 import synthetic
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = synthetic.get_workload(args.workload)
    res = cpu_baseline(w, steps=args.steps, warmup=args.warmup, cores=cores, max_seconds=240.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w),
        "cpu_baseline": {"value": res["value"], "unit": "pairs/s", "cores": cores, "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def torch_cuda_baseline(w, dev, model_ours, iters=10):
    """north_star's '>= 10x the reference PyTorch-CUDA forward on 1xB200': the UNMODIFIED reference modules on the GPU
    (torch eager, cuBLAS) next to ours, same batches. forward = eval mode, no_grad; fwd_bwd = the training step."""
    import torch
    from get_b200 import synthetic
    from get_b200.evaluate import CapturedForward
    if reference_modules() is None:
        return {"unavailable": "oracle/_ref not present (run python oracle/make_ref.py in the build container)"}

    def timed(fn, n):
        for i in range(4):                 # every one of the 4 batch shapes once (our side captures a graph per shape)
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    step, forward, pairs_list, ref_model = reference_step_fn(w, str(dev), train=True)
    ref_fb = timed(step, iters)
    ref_model.train(False)
    ref_fwd = timed(forward, iters)
    del ref_model
    batches = make_batches(w, 4, 123756)
    tens = [synthetic.batch_to_torch(b, device=dev) for b in batches]
    was = model_ours.training
    model_ours.train(False)
    cap = CapturedForward(model_ours)
    ours_fwd = timed(lambda i: cap.forward(tens[i % 4][0], tens[i % 4][1], tens[i % 4][3]), 4 * iters)
    model_ours.train(was)
    return {"reference_forward_ms": ref_fwd, "reference_fwd_bwd_adam_ms": ref_fb, "ours_forward_ms": ours_fwd,
            "forward_speedup": ref_fwd / ours_fwd, "pairs_per_batch": float(np.mean(pairs_list)),
            "note": "unmodified reference modules (.cuda(), torch eager) vs ours (captured forward), eval-mode forward over the "
                    "same 4 batches; fwd_bwd = zero_grad + forward + CE + backward + Adam in train mode"}


def graph_timed_ms(fn, reps=3):
    """Device time of fn() (a sequence of launches on the current stream) with the host out of the picture: the launches
    are captured ONCE into a CUDA graph and the graph replay is timed between two events (median of `reps` replays after
    one warm-up replay). A Python / ctypes launch costs ~10 us of host time -- more than some of the kernels measured."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()                                    # warm-up outside capture (lazy module loads, attribute opt-ins)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        g.replay()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
    torch.cuda.current_stream().wait_stream(side)
    del g
    return float(np.median(ms))


def stream_roofline(w, dev, graphs=7680, iters=10, sweep=False):
    """The fused GSL kernel at streaming size (BASELINE.json configs[4]: thousands of claim-evidence graphs resident, Snopes
    dims; SURVEY.md 8(d) asks for the roofline at B1 >= 1e4-ish sizes too). Inputs (2 sets, alternated) exceed L2.
    sweep=True: gnn_window_size in {3,5,9} x gsl_rate in {0.3,0.6,0.9} (train-mode launches)."""
    import torch
    from get_b200 import ops, synthetic
    N, H = w.len_right, w.hidden
    peak, _ = _peaks()
    wp, gate = torch.randn(H, device=dev) * 0.1, torch.randn(12, device=dev)
    npl = ops.gemm_mode(False)

    def adj_set(window):
        rng = np.random.default_rng(7 + window)
        base = []
        for g in range(48):
            pool = rng.integers(2, w.vocab, size=min(w.evd_pool, w.vocab - 2))
            toks = pool[rng.integers(0, pool.shape[0], size=N)]
            base.append(synthetic.word_graph(toks, N, window)[1].astype(np.float32))
        a1 = torch.from_numpy(np.stack(base)).to(dev)
        return a1.repeat((graphs + 47) // 48, 1, 1)[:graphs].contiguous(), float(np.mean([(b != 0).sum() for b in base]))

    feats = [torch.randn(graphs, N, H, device=dev) for _ in range(2)]

    def run(adj, k, p):
        # the scorer projection arrives precomputed, as in the model (by-product of the GEMM that writes the features), and
        # so do the neighbour lists (built once per step, shared by five consumers)
        sps = [ops.rowdot(f.view(graphs * N, H), wp, p, 1) for f in feats]
        lists = ops.NeighborLists(adj)
        recs = []
        ops.PROFILE_GSL_ARGS = recs
        for i in range(2):
            ops.gsl_fused(lists, feats[i], wp, gate, k, drop_p=p, seed_scorer=1, seed_layer2=2, planes_n=npl, sp_parts=sps[i])
        ops.PROFILE_GSL_ARGS = None
        return graph_timed_ms(lambda: [ops.gsl_fused_replay(recs[i % 2]) for i in range(iters)]) / iters

    class _W(object):
        len_right, hidden = N, H

    def own_bytes(nnz):
        return float(gsl_kernel_own_bytes(_W, graphs, nnz, 1, npl))

    out = {"graphs": graphs, "peak": peak, "unit": "GB/s",
           "bytes": "kernel-own bytes (F1 read + aggregation written + lists + scorer partial sums), see roofline.bytes",
           "timing": "10 launches captured in one CUDA graph, alternating two input sets (2 x 0.9 GB > L2)"}
    adj, nnz = adj_set(w.window)
    nbytes = own_bytes(nnz)
    out["algorithmic_bytes_per_launch"] = nbytes
    for name, p in (("eval", 0.0), ("train_p0.2", 0.2)):
        ms = run(adj, int(w.gsl_rate * N), p)
        out[name] = {"avg_launch_ms": ms, "achieved": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    lists = ops.NeighborLists(adj)
    bms = graph_timed_ms(lambda: [lists.rebuild() for _ in range(4)]) / 4
    dense_bytes = graphs * (4 * N * N + 2 * (8 * nnz + 4 * N))
    out["list_build"] = {"avg_launch_ms": bms, "bytes": dense_bytes, "achieved": dense_bytes / bms / 1e6, "frac": dense_bytes / bms / 1e6 / peak}
    del lists
    if sweep:
        rows = []
        for window in (3, 5, 9):
            adj, nnz = adj_set(window)
            nb = own_bytes(nnz)
            for rate in (0.3, 0.6, 0.9):
                ms = run(adj, int(rate * N), 0.2)
                rows.append({"gnn_window_size": window, "gsl_rate": rate, "nnz_per_graph": nnz, "avg_launch_ms": ms,
                             "algorithmic_bytes_per_launch": nb, "achieved": nb / ms / 1e6, "frac": nb / ms / 1e6 / peak})
        out["sweep_train"] = rows
    return out


def build_trainer(w, dev, precision, use_graph=True, flat_adam=True, micro=False):
    """Model + gradient bucket (the gradient storage) + Adam + captured step for workload w."""
    import torch
    from get_b200 import ops, synthetic
    from get_b200.ddp import FlatAdam, FlatGradAllReduce, trainable_named_parameters
    from get_b200.model import Graph_basedSemantiStructure
    from get_b200.step_graph import CapturedTrainStep
    ops.set_precision(precision)
    torch.manual_seed(123756)
    model = Graph_basedSemantiStructure(synthetic.match_params(w, cuda=True)).to(dev)
    model.train()
    named = trainable_named_parameters(model)
    reducer = FlatGradAllReduce([p for _, p in named], names=[n for n, _ in named])
    if flat_adam:
        opt = FlatAdam(reducer, lr=1e-4, weight_decay=1e-3)
    else:
        opt = torch.optim.Adam([p for _, p in named], lr=1e-4, weight_decay=1e-3, fused=True, capturable=use_graph)
    stepper = None
    if use_graph:
        stepper = CapturedTrainStep(model, None if micro else opt, reducer, accumulate=micro,
                                    collective_in_graph=os.environ.get("GET_B200_GRAPH_COLLECTIVE", "1") != "0")
    else:
        reducer.attach()
    return model, reducer, opt, stepper


def extra_workload(name, dev, precision, steps, warmup):
    """Short single-GPU measurement of another BASELINE.json config with the same step definition (extra keys of the line)."""
    import torch
    from get_b200 import _lib, ops, synthetic
    from get_b200.step_graph import pad_batch
    w = synthetic.get_workload(name)
    micro = name == "synthetic512"
    if micro:
        # configs[3]: B=512 claims x 30 evidences (15 360 pairs), R=200, D=H=512, 8 word heads. One optimizer step =
        # 16 micro-batches of 32 claims whose gradients accumulate in the bucket (the activations of all 3.07 M graph
        # rows at once would not fit), then one Adam step. The micro-batches alternate between 2 distinct synthetic ones.
        mb = 32
        model, reducer, opt, stepper = build_trainer(w, dev, precision, micro=True)
        batches = make_batches(w, 2, 777, n_claims=mb)
        tens = [synthetic.batch_to_torch(b, device=dev, adj_dtype=torch.float32) for b in batches]
        n_micro = w.batch_claims // mb
        stepper.loss_scale = 1.0 / n_micro

        def step(i):
            reducer.zero()
            for j in range(n_micro):
                loss = stepper.step(*tens[(i + j) % 2], mb)
            reducer.reduce()
            opt.step()
            return loss
        pairs_per_step = float(np.mean([b["pairs"] for b in batches])) * n_micro
    else:
        model, reducer, opt, stepper = build_trainer(w, dev, precision)
        batches = [pad_batch(b, PAD_PAIRS) for b in make_batches(w, 8, 4242)]
        tens = [synthetic.batch_to_torch(b, device=dev) for b in batches]

        def step(i):
            b = batches[i % 8]
            return stepper.step(*tens[i % 8], b["n_real_claims"])
        pairs_per_step = float(np.mean([b["real_pairs"] for b in batches]))
    for i in range(max(warmup, 8 if not micro else 1)):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"config": config_dict(w), "precision": precision, "dtype": "bf16" if precision == "bf16" else "f32", "steps": steps,
           "ms_per_step": ms, "value": pairs_per_step / ms * 1e3, "unit": "pairs/s", "pairs_per_step": pairs_per_step,
           "loss_finite": bool(torch.isfinite(loss).item()), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if micro:
        out["micro_batches"] = "%d x %d claims per optimizer step" % (w.batch_claims // 32, 32)
    reducer.detach()
    del model, reducer, opt, stepper, tens
    torch.cuda.empty_cache()
    return out


# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from get_b200 import _lib, ops, synthetic
    from get_b200.ddp import shard_claims
    from get_b200.keywords import KeyWordSettings as K

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...") and logs go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    w = synthetic.get_workload(args.workload)
    use_graph = not args.no_graph
    if args.workload == "synthetic512":
        if rank == 0:
            res = extra_workload("synthetic512", dev, args.precision, max(1, min(args.steps, 3)), 1)
            emit({"metric": METRIC, "value": res["value"], "unit": "pairs/s", "n_gpus": 1, "steps": res["steps"], "warmup": 1,
                  "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                  "dtype": res["dtype"], "data": "synthetic", "config": res["config"], "detail": res})
        return
    model, reducer, opt, stepper = build_trainer(w, dev, args.precision, use_graph=use_graph, flat_adam=not args.torch_adam)

    # every rank generates the same GLOBAL batches (world x 32 claims) and takes its claim shard, balanced by evidence count
    from get_b200.ddp import balance_claims
    from get_b200.step_graph import pad_batch, select_claims
    glob = make_batches(w, NBATCH, 123756, n_claims=w.batch_claims * world)
    if world > 1:
        batches = [select_claims(g, balance_claims(g[K.EvidenceCountPerQuery], world)[rank]) for g in glob]
    else:
        batches = glob[:]
    gclaims = [g["query"].shape[0] if world > 1 else 0 for g in glob]
    del glob
    padded = [pad_batch(b, PAD_PAIRS) for b in batches] if use_graph else batches
    n_real = [b.get("n_real_claims", b["query"].shape[0]) for b in padded]
    host = [synthetic.batch_to_torch(b, device="cpu", pin=True) for b in padded]
    resident = [synthetic.batch_to_torch(b, device=dev) for b in padded]

    def to_device(hb):
        q, d, l, kw = hb
        mv = lambda t: t.to(dev, non_blocking=True) if torch.is_tensor(t) else t
        kw2 = {k: (tuple(mv(x) for x in v) if isinstance(v, tuple) else mv(v)) for k, v in kw.items()}
        return mv(q), mv(d), mv(l), kw2

    def h2d_bytes(hb):
        q, d, l, kw = hb
        n = q.numel() * q.element_size() + d.numel() * d.element_size() + l.numel() * l.element_size()
        for v in kw.values():
            for x in (v if isinstance(v, tuple) else (v,)):
                if torch.is_tensor(x):
                    n += x.numel() * x.element_size()
        return n

    def eager_step(db, b):
        q, d, l, kw = db
        reducer.zero()
        logits = model(q, d, **kw)
        loss = ops.cross_entropy(logits[:n_real[b]], l[:n_real[b]])
        loss.backward()
        reducer.reduce()
        opt.step()
        return loss

    def step(db, b):
        if use_graph:
            return stepper.step(*db, n_real[b], global_claims=gclaims[b])      # copies into static buffers, one replay
        return eager_step(db, b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ---------------------------------------------------------------------------------
    for i in range(max(3, args.warmup)):
        step(resident[i % NBATCH], i % NBATCH)
    if use_graph:                      # build every graph before the timed regions (one per distinct padded shape)
        for b in range(NBATCH):
            step(resident[b], b)
    barrier()

    # ---- timed region 1: inputs resident in HBM --------------------------------------------------
    sampler = make_clock_sampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count() + (stepper.replayed_launches if use_graph else 0)
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pairs = 0
    for i in range(args.steps):
        b = (i + args.warmup) % NBATCH
        step(resident[b], b)
        pairs += batches[b]["pairs"]
    e1.record()
    t_enqueue = time.perf_counter() - t_wall0      # host time to enqueue the K steps (no sync inside)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() + (stepper.replayed_launches if use_graph else 0) - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = e0.elapsed_time(e1) / 1e3

    # ---- timed region 2: end to end from pinned host memory --------------------------------------
    def e2e_pass(nsteps):
        """nsteps end-to-end steps: per step the H2D copy of that step's inputs (overlapped with the previous step's
        compute on a side stream when graphs are on), the step, and a D2H read of the loss."""
        npairs, nh2d = 0, 0
        nxt = stepper.prefetch(*host[args.warmup % NBATCH]) if use_graph else None
        for i in range(nsteps):
            b = (i + args.warmup) % NBATCH
            if use_graph:
                loss = stepper.step_prefetched(nxt, n_real[b], gclaims[b])    # device-to-device copy-in + one graph replay
                if i + 1 < nsteps:                                         # H2D of the next batch while this step runs
                    nxt = stepper.prefetch(*host[(i + 1 + args.warmup) % NBATCH])
            else:
                loss = step(to_device(host[b]), b)
            _ = float(loss.item())                      # device -> host read of the step's result
            npairs += batches[b]["pairs"]
            nh2d += h2d_bytes(host[b])
        return npairs, nh2d

    e2e_pass(NBATCH)                                    # untimed: staging buffers of every batch shape get allocated here
    # three timed passes of K steps each; the MEDIAN pass is reported and all three are listed: on the shared hosts of this
    # pool single passes of this loop (a host sync every step) were observed to vary by tens of percent run to run
    e2e_times = []
    for _ in range(3):
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record()
        pairs_e2e, h2d = e2e_pass(args.steps)
        e3.record()
        barrier()
        e2e_times.append(max(time.perf_counter() - t0, e2.elapsed_time(e3) / 1e3))
    d2h = 4 * args.steps
    t_e2e = sorted(e2e_times)[1]

    # ---- diagnostic: what the pinned host -> device path delivers on this (shared) host right now ---------------
    h2d_diag = None
    if rank == 0:
        big = max((t for t in host[0][3].values() if torch.is_tensor(t)), key=lambda t: t.numel() * t.element_size())
        dst = torch.empty(big.shape, dtype=big.dtype, device=dev)
        rates = []
        for _ in range(8):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(big, non_blocking=True)
            b_.record()
            torch.cuda.synchronize()
            rates.append(big.numel() * big.element_size() / (a.elapsed_time(b_) * 1e-3) / 1e9)
        h2d_diag = {"bytes": big.numel() * big.element_size(), "pinned": bool(big.is_pinned()),
                    "GBs_min": min(rates), "GBs_median": float(np.median(rates)), "GBs_max": max(rates)}

    # ---- timed region 3: end to end from the COMPACT host format (token ids instead of dense float64
    #      adjacencies; the word graphs are built on the device by get_build_word_graphs) -----------------------------
    e2e_tok = None
    if use_graph and world == 1:
        from get_b200.step_graph import device_batch_from_tokens, token_batch_to_host
        tbs = [token_batch_to_host(b) for b in padded]
        tok_bytes = [sum(v.numel() * v.element_size() for v in tb.values() if torch.is_tensor(v)) for tb in tbs]
        for rep in range(2):                      # graphs for the fp32-adjacency input signature, then an allocator warm-up pass
            for b in range(NBATCH):
                float(stepper.step(*device_batch_from_tokens(tbs[b], dev), n_real[b]).item())
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e4.record()
        pairs_tok, h2d_tok = 0, 0
        nxt = stepper.prefetch_tokens(tbs[args.warmup % NBATCH])
        for i in range(args.steps):
            b = (i + args.warmup) % NBATCH
            loss = stepper.step_prefetched(nxt, n_real[b])
            if i + 1 < args.steps:            # token H2D + graph construction of the next batch while this step runs
                nxt = stepper.prefetch_tokens(tbs[(i + 1 + args.warmup) % NBATCH])
            _ = float(loss.item())
            pairs_tok += batches[b]["pairs"]
            h2d_tok += tok_bytes[b]
        e5.record()
        barrier()
        t_tok = max(time.perf_counter() - t0, e4.elapsed_time(e5) / 1e3)
        e2e_tok = {"value": pairs_tok / t_tok, "unit": "pairs/s (this rank)", "h2d_bytes_per_step": h2d_tok // args.steps,
                   "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * t_tok / args.steps,
                   "note": "host sends raw token ids; node lists and normalised adjacencies are built on the GPU, on a side stream "
                           "while the previous step runs (CapturedTrainStep.prefetch_tokens)"}

    # ---- roofline of the fused GSL kernel: the same steps issued kernel by kernel (events cannot be read out of a graph
    #      replay), CUDA events on the launch stream around every fused GSL launch --------------------------------------
    gsl_ms, gsl_pairs, stream_roof, tcb, extra = [], [], None, None, {}
    if rank == 0:
        # (1) the same steps issued kernel by kernel, recording the arguments of every fused GSL launch
        ops.PROFILE_GSL_ARGS = []
        reducer.overlap = False
        for i in range(min(args.steps, NBATCH)):
            b = (i + args.warmup) % NBATCH
            q, d, l, kw = resident[b]
            reducer.zero()
            loss = ops.cross_entropy(model(q, d, **kw)[:n_real[b]], l[:n_real[b]])
            loss.backward()
        torch.cuda.synchronize()
        recs, ops.PROFILE_GSL_ARGS = ops.PROFILE_GSL_ARGS, None
        # (2) exactly those launches (same neighbour lists / layer-1 features / dropout seeds, ~55 MB of distinct inputs and
        #     outputs each, ~1 GB in total > L2) captured into one CUDA graph and replayed between two events: the GPU never
        #     waits for the host, so the event interval is kernel time.
        ngraphs = [r[1].shape[0] for r in recs]
        for rep in range(2):
            gsl_ms.append(graph_timed_ms(lambda: [ops.gsl_fused_replay(r) for r in recs]) / len(recs))
            gsl_pairs.append(float(np.mean(ngraphs)))
        gsl_mode = recs[0][12]
        gsl_nnz = float(np.mean([float(r[0].nnz().sum().item()) / r[1].shape[0] for r in recs])) if gsl_mode == "lists" else None
        gsl_nsp = int(recs[0][11].shape[0]) if recs[0][11] is not None else 0
        gsl_planes = recs[0][10].nplanes if hasattr(recs[0][10], "nplanes") else 0
        build_ms = (graph_timed_ms(lambda: [r[0].rebuild() for r in recs]) / len(recs)) if gsl_mode == "lists" else 0.0
        real_pairs = float(np.mean([batches[(i + args.warmup) % NBATCH]["pairs"] for i in range(min(args.steps, NBATCH))]))
        del recs
        if world == 1 and not args.quick:
            stream_roof = stream_roofline(w, dev, sweep=True)
            tcb = torch_cuda_baseline(w, dev, model)
    barrier()

    # ---- reduce over ranks: max time, summed pairs ------------------------------------------------
    stats = torch.tensor([t_dev, t_e2e, float(pairs), float(pairs_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = stats.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = stats.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        rank_spread = {"t_dev_max_s": float(tmax[0]), "t_dev_min_s": float(tmin[0]), "pairs_max": float(tmax[2]), "pairs_min": float(tmin[2])}
        t_dev, t_e2e = float(tmax[0]), float(tmax[1])
        pairs, pairs_e2e = float(tsum[2]), float(tsum[3])
    else:
        rank_spread = None

    stepper_graphs = stepper.n_graphs() if use_graph else 0
    if rank == 0:
        peak, peak_src = _peaks()
        gsl_avg_ms = float(np.mean(gsl_ms)) if gsl_ms else None
        # algorithmic bytes on the REAL pairs of the batches (the launches also carry the dummy pairs that pad the batch to a
        # multiple of 16: they are work of the launch but not of the workload)
        survey_bytes = float(graph_kernel_bytes(w, real_pairs)) if gsl_ms else None
        if gsl_ms and gsl_mode == "lists":
            gsl_bytes = float(gsl_kernel_own_bytes(w, real_pairs, gsl_nnz, gsl_nsp, gsl_planes))
        else:
            gsl_bytes = survey_bytes
        achieved = gsl_bytes / (gsl_avg_ms * 1e-3) / 1e9 if gsl_ms else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gsl_fused_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        cpu = None
        if world == 1:                 # the CPU baseline is an N=1 measurement (other ranks' host threads would disturb it)
            cores = os.cpu_count() or 1
            cpu = cpu_baseline(w, steps=5, warmup=1, cores=cores, max_seconds=25.0)
            if not args.quick and args.workload == "snopes":
                # the other BASELINE.json configs, same step definition, short single-GPU runs
                reducer.detach()
                del stepper, opt
                torch.cuda.empty_cache()
                for name, prec, k in (("politifact", "bf16", 20), ("politifact", "fp32", 20), ("synthetic512", "fp32", 2)):
                    try:
                        extra["%s_%s" % (name, prec)] = extra_workload(name, dev, prec, k, 3)
                    except Exception as e:  # noqa: BLE001 -- an extra must never take the headline line down
                        extra["%s_%s" % (name, prec)] = {"error": str(e).splitlines()[-1][:300]}
                ops.set_precision(args.precision)
        line = {
            "metric": METRIC, "value": pairs / t_dev, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": config_dict(w),
            "run": {"parallelism": "dp%d" % world, "precision": args.precision,
                    "precision_note": {"fp32": "16-bit bf16-plane operands (3 tensor-core products), fp32-exact class (6 products) on the GSL top-k chain",
                                       "fp32x": "fp32-exact class everywhere", "bf16": "plain bf16 operands, fp32-exact class on the GSL top-k chain"}[args.precision],
                    "mean_pairs_per_step_per_gpu": pairs / args.steps / world,
                    "grad_allreduce_bytes": reducer.nbytes if world > 1 else 0, "grad_allreduce": (("one all-reduce after the backward pass (GET_B200_NO_OVERLAP=1)" if os.environ.get("GET_B200_NO_OVERLAP") == "1"
                                                                 else "3 chunks overlapped with the backward pass") if world > 1 else None),
                    "claims_sharding": ("global batch of %d claims per step, claims dealt to ranks by decreasing evidence count (least-loaded rank first)"
                                        % (w.batch_claims * world)) if world > 1 else None,
                    "rank_spread": rank_spread, "wall_s_timed_region": t_wall, "host_enqueue_ms_per_step": 1e3 * t_enqueue / args.steps,
                    "optimizer": "torch.optim.Adam(fused)" if args.torch_adam else "get_adam_flat_f32 (one kernel over the flat bucket)"},
            "clocks": clocks,
            "e2e": {"value": pairs_e2e / t_e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "passes_ms_per_step": [1e3 * t / args.steps for t in e2e_times], "reported": "median of 3 passes of K steps"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "fused GSL graph kernel (get_gsl_gather: scorer + top-k + refined aggregation on neighbour lists), train-mode "
                                   "dropout, the bench batches' own launches replayed from one CUDA graph",
                         "bound": "hbm", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "avg_launch_ms": gsl_avg_ms, "algorithmic_bytes_per_launch": gsl_bytes,
                         "bytes": "F1 read + refined aggregation written + neighbour lists + scorer partial sums (the dense adjacency is read "
                                  "by the list builder, once per step for five consumers)",
                         "list_build_ms": build_ms if gsl_ms else None,
                         "survey_formula": ({"bytes_per_launch": survey_bytes, "note": "4R(2H+R) per pair incl. the dense adjacency, over fused kernel + list build time",
                                             "frac": survey_bytes / ((gsl_avg_ms + build_ms) * 1e-3) / 1e9 / peak} if gsl_ms else None),
                         "bytes_counted_on": "real pairs (%.1f per launch; launches carry %.1f incl. padding)" % (real_pairs, float(np.mean(gsl_pairs)) if gsl_pairs else 0.0),
                         "frac_of_8TBs_nominal": (achieved / 8000.0) if achieved else None},
            "e2e_token_inputs": e2e_tok,
            "h2d_diagnostic": h2d_diag,
            "roofline_stream": stream_roof,
            "torch_cuda_baseline": tcb,
            "extra": extra or None,
            "cuda_graphs": {"enabled": use_graph, "graphs": stepper_graphs, "pad_pairs_to": PAD_PAIRS if use_graph else 0},
            "cpu_baseline": ({"value": cpu["value"], "unit": "pairs/s", "cores": os.cpu_count(), "kind": cpu["kind"],
                              "sample": cpu["sample"], "ms_per_step": cpu["ms_per_step"]} if cpu is not None else None),
        }
        emit(line)
    if world > 1:
        # Tearing down a communicator whose collectives live inside captured CUDA graphs can block in ncclCommDestroy;
        # every rank is done once it passes this barrier, so leave without destroying the process group.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=["snopes", "politifact", "synthetic512", "stream"],
                    help="BASELINE.json configs: snopes (configs[0]/[1], the headline), politifact (configs[2]), synthetic512 "
                         "(configs[3], micro-batched), stream (configs[4] shape)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp32x", "bf16", "fast"],
                    help="fp32 = 16-bit plane operands outside the GSL top-k chain (the judged 1e-4 configuration); fp32x = "
                         "fp32-exact class everywhere; bf16 = plain bf16 operands outside the chain (1e-2 class)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam(fused) instead of the in-tree flat Adam kernel")
    ap.add_argument("--quick", action="store_true", help="skip the extras (other workloads, stream sweep, torch-CUDA baseline)")
    args = ap.parse_args()
    if args.precision == "fast":
        args.precision = "bf16"
    guard_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
