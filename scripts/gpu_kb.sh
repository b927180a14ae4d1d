#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py tests/test_gpu_tc.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -8 | cut -c1-400
timeout 300 python bench.py --precision fast --steps 30 --warmup 5 > $OUT/bench_fast.json 2> $OUT/bench_fast.err; tail -2 $OUT/bench_fast.err | cut -c1-300; cut -c1-900 $OUT/bench_fast.json
