python -m pytest tests/test_gpu_clam.py -q -p no:cacheprovider 2>&1 | tail -3
CFG_REPS=6 CFG_CPU=0 python tools/bench_configs.py 2>/dev/null | tail -n 7
