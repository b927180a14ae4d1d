#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|near-tie|torch_cuda|speedup" | head -30 | cut -c1-600
cat $OUT/torch_cuda_baseline.json 2>/dev/null
python scripts/kbench.py graph --G 216 2>&1 | tail -1
python scripts/kbench.py graph --G 7680 2>&1 | tail -1
GET_B200_GRAPH_CLUSTER=0 python scripts/kbench.py graph --G 216 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_c.json 2> $OUT/bench_c.err; tail -3 $OUT/bench_c.err | cut -c1-300; cat $OUT/bench_c.json | cut -c1-2600
