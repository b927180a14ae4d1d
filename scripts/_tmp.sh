timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short -p no:cacheprovider -k segments 2>&1 | grep -E "^E  |tests/test_gpu_tc.py:" | head -5 | cut -c1-200
