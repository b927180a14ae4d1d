#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
M=gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__occupancy_limit_warps,sm__warps_active.avg.pct_of_peak_sustained_active,launch__shared_mem_per_block_dynamic,launch__shared_mem_config_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k "regex:gather_row_kernel|build_neighbor" -c 40 --csv --log-file $OUT/occ.csv python scripts/dbg_graph.py > $OUT/occ.log 2>&1
python - <<'PY'
import csv, collections
rows=[l for l in open("gpurun_out/occ.csv") if not l.startswith("==")]
d=collections.OrderedDict()
for r in csv.DictReader(rows):
    k=(r["ID"], r["Kernel Name"][:60], r["Grid Size"])
    d.setdefault(k,{})[r["Metric Name"].split("__")[-1][:28]]=r["Metric Value"]
for k,v in d.items():
    print(k, " ".join("%s=%s"%x for x in v.items()))
PY
