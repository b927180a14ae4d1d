#!/bin/bash
# all GPU parity tests + smoke; prints the tail
TAG=${1:-t}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider > $OUT/tests_$TAG.log 2>&1
tail -${2:-60} $OUT/tests_$TAG.log | cut -c1-260
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
