#!/bin/bash
OUT=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_r1c.csv \
    python bench.py --no-graph --steps 2 --warmup 3 > $OUT/ncu_launch_r1c.log 2>&1
tail -1 $OUT/ncu_launch_r1c.log | cut -c1-200
