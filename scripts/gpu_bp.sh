#!/bin/bash
# bring-up of the plane GEMM: probe, kernel tests, then the whole GPU suite and a bench line
TAG=${1:-bp}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/bp_probe.py > $OUT/probe_$TAG.log 2>&1; cat $OUT/probe_$TAG.log | grep -v "^  File\|^Traceback\|^    " | tail -30
timeout 600 python -m pytest tests/test_gpu_gemm_bp.py tests/test_gpu_tc.py -q --maxfail=40 --tb=line -p no:cacheprovider > $OUT/tests_bp_$TAG.log 2>&1
tail -45 $OUT/tests_bp_$TAG.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider --deselect tests/test_gpu_gemm_bp.py --deselect tests/test_gpu_tc.py > $OUT/tests_$TAG.log 2>&1
tail -40 $OUT/tests_$TAG.log | cut -c1-250
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json | cut -c1-1500; tail -5 $OUT/bench_$TAG.err
