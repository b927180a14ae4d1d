#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -15 | cut -c1-250
