#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_lists.py -q --maxfail=20 --tb=short -p no:cacheprovider > $OUT/tests_lists.log 2>&1
tail -30 $OUT/tests_lists.log | cut -c1-250
timeout 300 python scripts/dbg_graph.py > $OUT/dbg_graph.log 2>&1; grep TIME $OUT/dbg_graph.log; grep -v TIME $OUT/dbg_graph.log | tail -10
timeout 300 python scripts/dbg_graph_cold.py > $OUT/dbg_graph_cold.log 2>&1; grep TIME $OUT/dbg_graph_cold.log; grep -v TIME $OUT/dbg_graph_cold.log | tail -10
timeout 600 python bench.py --quick --steps 20 --warmup 5 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; python -c "
import json; d=json.load(open('$OUT/bench_quick.json')); print('ms_per_step', d['ms_per_step'], 'roofline', {k: d['roofline'][k] for k in ('frac','avg_launch_ms','list_build_ms','algorithmic_bytes_per_launch')}, d['roofline']['survey_formula'])"; tail -3 $OUT/bench_quick.err | cut -c1-300
