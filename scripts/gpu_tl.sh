#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/dbg_graph.py > $OUT/dbg_graph_tl.log 2>&1
grep TIME $OUT/dbg_graph_tl.log | head -8
for f in 0 1; do for c in 0 100 219 3000; do echo "--- fused $f cta $c"; grep "GRDBG cta $c fused $f" $OUT/dbg_graph_tl.log | awk 'NR%5==2' | head -8; done; done
