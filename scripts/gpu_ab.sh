#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2; do timeout 300 python bench.py --quick --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('run $i ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"; done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/ab_step_launches.csv python scripts/one_step.py snopes fp32 > $OUT/ab_step.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[l for l in open("gpurun_out/ab_step_launches.csv") if not l.startswith("==")]
agg=collections.OrderedDict(); tot=0
for r in csv.DictReader(rows):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","")[:60]
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print("total us %.1f"%tot)
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print("%-62s %4d %9.1f %5.1f%% %7.1f"%(k,c,t,100*t/tot,t/c))
PY
