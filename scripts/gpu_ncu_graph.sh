#!/bin/bash
# ncu --set full of the graph kernels as launched by scripts/dbg_graph.py (B=32-sized launches come first)
OUT=gpurun_out
mkdir -p $OUT
K=${1:-gather_row_kernel}
SKIP=${2:-20}
CNT=${3:-3}
TAG=${4:-x}
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s $SKIP -c $CNT -f -o $OUT/prof_$TAG \
   python scripts/dbg_graph.py > $OUT/ncu_$TAG.log 2>&1
ls -la $OUT/prof_$TAG.ncu-rep; tail -3 $OUT/ncu_$TAG.log
