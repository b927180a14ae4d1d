#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $OUT/tests_tc.log 2>&1
echo "tc tests exit $?"; tail -25 $OUT/tests_tc.log
