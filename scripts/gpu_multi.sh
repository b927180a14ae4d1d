#!/bin/bash
# Multi-GPU sanity: our arm and the reference arm launched the way the driver launches them. Usage: bash scripts/gpu_multi.sh N
OUT=gpurun_out
N=${1:-2}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "exit $? lines $(wc -l < $OUT/bench_n$N.json)"; head -c 400 $OUT/bench_n$N.json; echo
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | cut -c1-200
