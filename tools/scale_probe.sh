python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
CFG_REPS=6 CFG_CPU=0 python tools/bench_configs.py 2>/dev/null | tail -n 48 | grep -E "\"teacher|forward_test|graphed|train_step_ms|cfg0" | head -24
for b in attn dsmil; do d=1024; [ $b = dsmil ] && d=1536; T_BASE=$b T_D=$d timeout 300 python tools/bench_train_step.py 2>/dev/null | tail -1; done
