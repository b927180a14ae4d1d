#!/bin/bash
# launch list of ONE training step (ncu durations, issue order) + optional full capture of one kernel family
TAG=${1:-p}
KREGEX=${2:-}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/step_$TAG.csv \
    python scripts/one_step.py snopes fp32 > $OUT/step_$TAG.log 2>&1
python - <<PY
import csv, collections, re
rows=[l for l in open("$OUT/step_$TAG.csv") if not l.startswith("==")]
agg=collections.OrderedDict(); tot=0
seq=[]
for r in csv.DictReader(rows):
    if r.get("Metric Name")!="gpu__time_duration.sum": continue
    name=re.sub(r"\(.*","",r["Kernel Name"]).replace("void ","")[:60]
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    v = v/1e3 if u=="ns" else (v*1e3 if u=="ms" else v)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v; seq.append((name,v,r.get("Grid Size","")))
print("total us %.1f launches %d"%(tot,len(seq)))
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:40]:
    print("%-62s %4d %9.1f %5.1f%% %7.1f"%(k,c,t,100*t/tot,t/c))
open("$OUT/step_${TAG}_seq.txt","w").write("\n".join("%-62s %8.1f %s"%s for s in seq))
PY
if [ -n "$KREGEX" ]; then
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$KREGEX -c 12 -f -o $OUT/prof_${TAG} \
    python scripts/one_step.py snopes fp32 > $OUT/ncu_${TAG}.log 2>&1
ls -la $OUT/prof_${TAG}.ncu-rep
fi
