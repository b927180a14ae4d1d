#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_gemm_bp.py tests/test_gpu_tc.py -q --maxfail=40 --tb=line -p no:cacheprovider > $OUT/tests_bp.log 2>&1
tail -15 $OUT/tests_bp.log | cut -c1-250
timeout 300 python scripts/dbg_bp.py > $OUT/dbg_bp.log 2>&1; grep TIME $OUT/dbg_bp.log; grep -v TIME $OUT/dbg_bp.log | tail -75
