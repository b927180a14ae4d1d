#!/bin/bash
for v in 1 1; do python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('run', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['e2e_token_inputs']['ms_per_step'],3), d['h2d_diagnostic'])"; done
