#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
python scripts/kbench.py accuracy 2>&1 | tail -2
python scripts/kbench.py graph --G 216 2>&1 | tail -1
python scripts/kbench.py graph --G 960 2>&1 | tail -1
python scripts/kbench.py graph --G 7680 2>&1 | tail -1
python scripts/kbench.py gemm --M 21600 --K 300 --seg 2 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_smem_kernel -s 4 -c 3 -f -o $OUT/prof_gs \
    python scripts/kbench.py graph --G 216 --iters 3 > $OUT/ncu_gs.log 2>&1
tail -3 $OUT/ncu_gs.log
