#!/bin/bash
OUT=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) bench.py --quick --gpus 2 --steps 30 --warmup 5 > $OUT/dbg_$name.json 2> $OUT/dbg_$name.err
  echo "== $name exit $? json bytes $(wc -c < $OUT/dbg_$name.json) $(python -c "import json;d=json.load(open('$OUT/dbg_$name.json'));print('ms/step',d['ms_per_step'])" 2>/dev/null)"; grep "rank[01]\]" $OUT/dbg_$name.err | grep "Error\|error" | grep -v "Warn" | tail -2 | cut -c1-200
}
run thread_local A=1
run global_mode GET_B200_CAPTURE_MODE=global
run graph_register0 GET_B200_CAPTURE_MODE=global NCCL_GRAPH_REGISTER=0
run nvls0 GET_B200_CAPTURE_MODE=global NCCL_NVLS_ENABLE=0
run split_tail GET_B200_CAPTURE_MODE=global GET_B200_SPLIT_TAIL=1
run ring_simple GET_B200_CAPTURE_MODE=global NCCL_ALGO=Ring NCCL_PROTO=Simple
