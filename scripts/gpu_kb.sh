#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | grep -E "^E  |passed|failed|FAILED|Error|near-tie|speedup" | head -12 | cut -c1-700
