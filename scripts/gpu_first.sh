#!/bin/bash
# First contact with the GPU: SIMT path first (GET_B200_TC=0), then the tcgen05 path under a short timeout.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpu_first.txt
GET_B200_TC=0 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $OUT/tests_simt.log 2>&1
tail -5 $OUT/tests_simt.log
GET_B200_TC=0 timeout 300 python __graft_entry__.py --smoke > $OUT/smoke_simt.log 2>&1; tail -2 $OUT/smoke_simt.log
GET_B200_TC=0 timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/bench_simt.json 2> $OUT/bench_simt.err; cat $OUT/bench_simt.json; tail -3 $OUT/bench_simt.err
timeout 240 python -m pytest tests/test_gpu_tc.py -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $OUT/tests_tc.log 2>&1
echo "tc tests exit $?"; tail -15 $OUT/tests_tc.log
nvidia-smi --query-gpu=name,memory.used --format=csv
