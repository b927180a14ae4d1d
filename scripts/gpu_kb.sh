#!/bin/bash
for K in 600; do for m in store; do echo "== K=$K $m"; GET_B200_T2_DEBUG=9 timeout 60 python scripts/dbg_t2.py $K $m 2>&1 | grep T2DBG | tail -8; done; done
