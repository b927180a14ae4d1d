#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/dbg_graph.py > $OUT/dbg_graph_tl.log 2>&1
grep TIME $OUT/dbg_graph_tl.log | head -8
for c in 0 100 219 3000; do echo "--- cta $c"; grep "GLDBG cta $c " $OUT/dbg_graph_tl.log | awk 'NR%13==5' | head -14; done
