#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/dbg_graph_cold.py > $OUT/dbg_graph_cold.log 2>&1
grep TIME $OUT/dbg_graph_cold.log | head -8
for c in 0 100 219; do echo "--- cta $c"; grep "GRDBG cta $c fused 1" $OUT/dbg_graph_cold.log | awk 'NR%9==4' | head -12; done
