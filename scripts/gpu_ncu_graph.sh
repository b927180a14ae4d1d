#!/bin/bash
# ncu --set full of the graph kernels as launched by scripts/dbg_graph.py (B=32-sized launches come first)
OUT=gpurun_out
mkdir -p $OUT
K=${1:-gather_row_kernel}
SKIP=${2:-20}
CNT=${3:-3}
GET_B200_GRAPH_KERNEL=${4:-row} timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o $OUT/prof_$K \
   python scripts/dbg_graph.py > $OUT/ncu_$K.log 2>&1
ls -la $OUT/prof_$K.ncu-rep; tail -3 $OUT/ncu_$K.log
