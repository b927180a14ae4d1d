#!/bin/bash
timeout 600 python -m pytest tests/test_step_graph.py tests/test_gpu_tc.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -E "^tests|^E  |passed|failed|Error|FAILED" | head -12 | cut -c1-400
