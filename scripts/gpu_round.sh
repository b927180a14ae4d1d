#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench line (ours + reference arm), ncu launch list and full captures.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpu_$TAG.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $OUT/tests_$TAG.log 2>&1
tail -3 $OUT/tests_$TAG.log
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke_$TAG.log 2>&1; tail -1 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; cat $OUT/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; cat $OUT/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_launch_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:graph_smem_kernel -s 40 -c 6 -f -o $OUT/prof_graph_$TAG \
    python bench.py --steps 1 --warmup 3 > $OUT/ncu_graph_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 300 -c 6 -f -o $OUT/prof_gemm_$TAG \
    python bench.py --steps 1 --warmup 3 > $OUT/ncu_gemm_$TAG.log 2>&1
ls -la $OUT | tail -12
