#!/bin/bash
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc2_kernel --csv --log-file $OUT/t_nt.csv python scripts/kbench.py gemm --M 21600 --N 300 --K 300 --seg 2 --iters 3 > /dev/null 2>&1
grep gemm_tc2 $OUT/t_nt.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -56 | tr '\n' ' '; echo
