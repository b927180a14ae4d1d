#!/bin/bash
# SASS evidence for the built library: tcgen05 / TMEM / TMA / mbarrier mnemonics per kernel (cuobjdump, no GPU needed).
SO=get_b200/csrc/libget_b200.so
OUT=${1:-profiles/r2_sass_summary.txt}
cuobjdump -sass $SO > /tmp/getb_sass.txt
python - "$OUT" <<'PY'
import collections, re, subprocess, sys
pat = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCATOM", "SYNCS", "LDGSTS", "HMMA", "FFMA", "LDS.128", "STG.E.128", "LDG.E.128", "F2FP.BF16"]
cur = None
per = collections.OrderedDict()
arch = None
for line in open("/tmp/getb_sass.txt"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for p in pat:
            if op.startswith(p):
                per[cur][p] += 1
with open(sys.argv[1], "w") as fh:
    fh.write("# cuobjdump -sass get_b200/csrc/libget_b200.so (%s): instruction counts per kernel -- UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM),\n"
             "# UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, UBLKPF = bulk prefetch, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit\n" % arch)
    tot = collections.Counter()
    fh.write("%-64s %7s  %s\n" % ("kernel", "instrs", "matched mnemonics"))
    for k, c in per.items():
        s = " ".join("%s=%d" % (p, c[p]) for p in pat if c[p])
        fh.write("%-64s %7d  %s\n" % (k[:64], c["_total"], s))
        tot.update(c)
    fh.write("\nTOTAL over %d kernels: %s\n" % (len(per), " ".join("%s=%d" % (p, tot[p]) for p in pat if tot[p])))
PY
head -5 $OUT; grep "gemm_bp_kernel<0>\|gather_row_kernel<true, 3, 2, true>\|TOTAL\|build_neighbor\|gather_row_kernel<(bool)1, (int)3, (int)2, (bool)1>" $OUT | cut -c1-300
