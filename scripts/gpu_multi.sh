#!/bin/bash
# Multi-GPU: NCCL alignment probe, 2-GPU gradient-equality test, then our arm at 2..N GPUs launched the way the driver
# launches it (with and without the overlapped all-reduce), plus the N=1 line on the same box. Usage: bash scripts/gpu_multi.sh N
OUT=gpurun_out
N=${1:-2}
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_trainer.py -m gpu -q --tb=short -p no:cacheprovider -k two_gpu > $OUT/tests_multi_n$N.log 2>&1; tail -3 $OUT/tests_multi_n$N.log | cut -c1-200
timeout 150 python bench.py --quick --steps 30 --warmup 5 > $OUT/bench_n1_same_box.json 2> $OUT/bench_n1_same_box.err
for M in 2 4 8; do
  if [ $M -le $N ]; then
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port $((29510+M)) bench.py --quick --gpus $M --steps 30 --warmup 5 > $OUT/bench_n$M.json 2> $OUT/bench_n$M.err
    echo "N=$M exit $? lines $(wc -l < $OUT/bench_n$M.json)"; grep -v "Warn\|warn\|^\*\|OMP_NUM" $OUT/bench_n$M.err | grep "Error\|error" | head -3 | cut -c1-300
    GET_B200_NO_OVERLAP=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port $((29520+M)) bench.py --quick --gpus $M --steps 30 --warmup 5 > $OUT/bench_n${M}_no_overlap.json 2> $OUT/bench_n${M}_no_overlap.err
  fi
done
python - <<'PY'
import json, glob, os
base = None
for f in ["gpurun_out/bench_n1_same_box.json"] + sorted(glob.glob("gpurun_out/bench_n[248]*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if base is None: base = d["value"]
    print("%-34s n=%d ms/step %.3f value %.0f eff %.3f | %s | spread %s" % (os.path.basename(f), d["n_gpus"], d["ms_per_step"], d["value"], d["value"] / d["n_gpus"] / base, d["run"].get("grad_allreduce"), d["run"].get("rank_spread")))
PY
