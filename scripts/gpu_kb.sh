#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -2 | cut -c1-250
for K in 32 300; do for m in store sig; do echo "== K=$K $m"; GET_B200_T2_DEBUG=9 python scripts/dbg_t2.py $K $m 2>&1 | grep T2DBG | tail -5; done; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_tc2_kernel --csv --log-file $OUT/t_d0.csv python scripts/kbench.py gemm --M 21600 --N 300 --K 300 --seg 2 --iters 3 > /dev/null 2>&1
echo "debug=0"; grep gemm_tc2 $OUT/t_d0.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -48 | tr '\n' ' '; echo
