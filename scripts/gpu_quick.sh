#!/bin/bash
# quick check after a kernel change: all GPU tests, then the short bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --tb=short -p no:cacheprovider > $OUT/tests_q.log 2>&1; grep -v "Warn\|warn" $OUT/tests_q.log | tail -12 | cut -c1-220
timeout 400 python bench.py --quick --steps 30 --warmup 5 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; python -c "
import json; d=json.load(open('$OUT/bench_quick.json')); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'tok', d['e2e_token_inputs']['ms_per_step'], 'roofline', {k: d['roofline'][k] for k in ('frac','avg_launch_ms','list_build_ms')})"; tail -2 $OUT/bench_quick.err | cut -c1-200
