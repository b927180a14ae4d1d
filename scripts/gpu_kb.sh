#!/bin/bash
OUT=gpurun_out
python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step',d['ms_per_step'],'enqueue',d['config']['host_enqueue_ms_per_step'],'e2e',d['e2e'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_t3.csv \
    python bench.py --steps 2 --warmup 3 > $OUT/ncu_launch_t3.log 2>&1
