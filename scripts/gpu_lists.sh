#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_lists.py -q --maxfail=20 --tb=short -p no:cacheprovider > $OUT/tests_lists.log 2>&1
tail -40 $OUT/tests_lists.log | cut -c1-250
timeout 300 python scripts/dbg_graph.py > $OUT/dbg_graph.log 2>&1; grep TIME $OUT/dbg_graph.log; grep -v TIME $OUT/dbg_graph.log | tail -20
timeout 600 python bench.py --quick --steps 20 --warmup 5 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; cat $OUT/bench_quick.json | cut -c1-1500; tail -5 $OUT/bench_quick.err
