timeout 900 python -m pytest tests/test_gpu_clam.py -x -q -p no:cacheprovider 2>&1 | tail -25
