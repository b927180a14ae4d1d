#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -8 | cut -c1-250
python scripts/kbench.py graph --G 216 2>&1 | tail -1
python scripts/kbench.py graph --G 960 2>&1 | tail -1
python scripts/kbench.py graph --G 7680 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:graph_smem_kernel -s 4 -c 2 -f -o $OUT/prof_gs3 \
    python scripts/kbench.py graph --G 216 --iters 3 > $OUT/ncu_gs3.log 2>&1
tail -2 $OUT/ncu_gs3.log
