#!/usr/bin/env python
"""Turn the raw ncu outputs of scripts/gpu_round.sh (gpurun_out/, scratch) into the small tracked summaries under
profiles/:  python scripts/summarize_profiles.py <tag>   (needs `ncu` for reading .ncu-rep files; no GPU)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(os.path.join(PROF, "%s_launches_summary.txt" % tag), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3\n")
        fh.write("# %d launches captured (graph-building steps are issued kernel by kernel, so every kernel of the step is\n"
                 "# listed); per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n" % len(rows))
        fh.write("%-78s %8s %12s %7s %9s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write("%-78s %8d %12.1f %6.1f%% %9.1f\n" % (k[:78], c, t, 100 * t / tot, t / c))
    return {k: (c, t / tot) for k, (c, t) in agg.items()}


def rep(tag, which, kernel_filter):
    path = os.path.join(OUT, "prof_%s_%s.ncu-rep" % (which, tag))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel_filter not in d.get("Kernel Name", ""):
            continue
        res.append({k: (d[k], units[hdr.index(k)]) for k in KEYS if k in d} | {"kernel": d["Kernel Name"]})
    with open(os.path.join(PROF, "%s_%s_ncu_full.txt" % (tag, which)), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:%s ... python bench.py --steps 1 --warmup 3\n"
                 % kernel_filter)
        for i, e in enumerate(res):
            fh.write("\n## launch %d: %s\n" % (i, e["kernel"]))
            for k in KEYS:
                if k in e:
                    fh.write("%-75s %s %s\n" % (k, e[k][0], e[k][1]))
    return res


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    shares = launches(tag)
    for name in ("bench_%s.json" % tag, "bench_ref_%s.json" % tag):
        src = os.path.join(OUT, name)
        if os.path.exists(src):
            with open(src) as fh, open(os.path.join(PROF, "%s_%s" % (tag, name.replace("_" + tag, ""))), "w") as out:
                out.write(fh.read())
    g = rep(tag, "graph", "graph_smem_kernel")
    fused = [e for e in g if "<(bool)1" in e["kernel"] or "<1" in e["kernel"]]
    if fused:
        def num(e, k):
            v, u = e[k]
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        tr = [num(e, "dram__bytes_read.sum") + num(e, "dram__bytes_write.sum") for e in fused]
        with open(os.path.join(PROF, "gsl_fused_traffic.json"), "w") as fh:
            json.dump({"source": "profiles/%s_graph_ncu_full.txt" % tag,
                       "kernel": "graph_smem_kernel<FUSED=1> (get_gsl_fused_f32), bench batch",
                       "dram_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr),
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch; writes partly stay in L2 under ncu"},
                      fh, indent=1)
    rep(tag, "gemm", "gemm_tc2_kernel")
    print("profiles written for", tag)


if __name__ == "__main__":
    main()
