#!/bin/bash
# quick GPU check: parity tests + bench line
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $OUT/tests_$TAG.log 2>&1
tail -12 $OUT/tests_$TAG.log | cut -c1-300
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
