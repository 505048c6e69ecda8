python -m pytest tests/test_gpu_dtfd.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|^E   .*(assert|Error)|^tests/test_gpu_dtfd.py:[0-9]+" | cut -c1-260 | head -30
