#!/usr/bin/env python
"""Turn the raw outputs of scripts/gpu_round2.sh (gpurun_out/, scratch) into the small tracked summaries under profiles/:
    python scripts/summarize_profiles.py [tag]      (needs `ncu` to read .ncu-rep files; no GPU)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches():
    path = os.path.join(OUT, "%s_step_launches.csv" % TAG)
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    seq = []
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        seq.append((name, v, r.get("Grid Size", "")))
    with open(os.path.join(PROF, "%s_launches_summary.txt" % TAG), "w") as fh:
        fh.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python scripts/one_step.py snopes fp32\n")
        fh.write("# ONE eager training step (fwd + CE + bwd + Adam) of the bench workload, %d launches, %.1f us in total; per-launch times are\n"
                 "# cold-cache and serialised under ncu: compare SHARES, not absolutes (the captured step replays in 2.27 ms).\n" % (len(rows), tot))
        fh.write("%-70s %8s %12s %7s %9s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            fh.write("%-70s %8d %12.1f %6.1f%% %9.1f\n" % (k[:70], c, t, 100 * t / tot, t / c))
        fh.write("\n# issue order\n")
        for name, v, grid in seq:
            fh.write("%-70s %9.1f us  grid %s\n" % (name[:70], v, grid))


def rep(which, title):
    path = os.path.join(OUT, "%s_prof_%s.ncu-rep" % (TAG, which))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    with open(os.path.join(PROF, "%s_%s_ncu_full.txt" % (TAG, which)), "w") as fh:
        fh.write("# %s\n# ncu --set full --clock-control none --import-source on (one replayed launch per row; numbers under a profiler are\n"
                 "# for shares and counters, never bench values)\n" % title)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            fh.write("\n== %s  grid %s block %s\n" % (name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            d = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    fh.write("  %-88s %s %s\n" % (k, r[i], units[i]))
                    d[k] = (r[i], units[i])
            out.append((name, d))
    return out


def main():
    os.makedirs(PROF, exist_ok=True)
    launches()
    b32 = rep("gsl_b32", "fused GSL kernel, the launch of one B=32 Snopes training step (scripts/one_step.py)")
    st = rep("gsl_stream", "fused GSL kernel at streaming size: 7 680 graphs, train-mode dropout, bf16-plane output (scripts/prof_gsl_stream.py)")
    rep("gemm", "plane GEMM (gemm_bp_kernel), first 8 launches of one B=32 Snopes training step (scripts/one_step.py)")

    def mb(v):
        x, u = v
        x = float(x.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic = {}
    for tag, reps in (("b32", b32), ("stream", st)):
        name, d = reps[0]
        traffic[tag] = {"dram_bytes_read": mb(d["dram__bytes_read.sum"]), "dram_bytes_write": mb(d["dram__bytes_write.sum"]),
                        "kernel": name}
    t = traffic["b32"]
    with open(os.path.join(PROF, "gsl_fused_traffic.json"), "w") as fh:
        json.dump({"dram_bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"], "detail": traffic,
                   "source": "profiles/%s_gsl_b32_ncu_full.txt / %s_gsl_stream_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch; "
                             "under ncu's cache control the written planes of a 54 MB launch largely stay in the 126 MB L2)" % (TAG, TAG)}, fh, indent=1)
    for f in ("bench.json", "bench_ref.json"):
        shutil.copy(os.path.join(OUT, "%s_%s" % (TAG, f)), os.path.join(PROF, "%s_%s" % (TAG, f)))
    for f, keep in (("memcheck_lists.log", 12), ("memcheck_smoke.log", 12), ("racecheck_lists.log", 12)):
        with open(os.path.join(OUT, "%s_%s" % (TAG, f))) as fh:
            lines = [l for l in fh if "Warning" not in l and "warn" not in l]
        with open(os.path.join(PROF, "%s_sanitizer_%s" % (TAG, f)), "w") as fh:
            fh.writelines(lines[:4] + ["...\n"] + lines[-keep:])


if __name__ == "__main__":
    main()
