"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout with the agreed
keys, and the compact token batch carries what the device-side graph construction needs."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    # the unmodified reference modules when oracle/_ref (or /root/reference) is present, else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    from oracle import ref_import
    assert (d["cpu_baseline"]["kind"] == "reference") == ref_import.available()
    assert set(d["config"].keys()) == {"workload", "step", "claims_per_gpu_per_step", "l2"}
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_token_batch_carries_the_raw_sequences():
    from get_b200 import synthetic
    from get_b200.keywords import KeyWordSettings as K
    from get_b200.step_graph import pad_batch, token_batch_to_host
    w = synthetic.get_workload("tiny")
    b = pad_batch(synthetic.make_batch(w, seed=4), 8)
    tb = token_batch_to_host(b, pin=False)
    assert tb["d_tok"].shape[0] == b["pairs"] == int(b[K.EvidenceCountPerQuery].sum())
    assert tb["q_tok"].shape[0] == b["query"].shape[0]
    # rebuilding the graphs on the host from the raw sequences reproduces the batch's node lists and adjacencies
    for g in range(b["pairs"]):
        L = int(tb["d_len"][g])
        nodes, adj, nn = synthetic.word_graph(tb["d_tok"][g, :L].numpy(), tb["R"], tb["window"])
        assert np.array_equal(nodes, b[K.DocContentNoPaddingEvidence][g]) and nn == int(b["e_lens"][g])
        assert np.array_equal(adj, b[K.Evd_Docs_Adj][g])
