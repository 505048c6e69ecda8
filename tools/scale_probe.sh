# start-up cost probe of the bench loop (see profiles/round2_summary.md): per-step GPU timestamps, then the bench line at K = 20 on 1 and 2 ranks
python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | grep -E "full-tensor gradient errors|passed|failed|^FAILED|^E  " | cut -c1-1500 | tail -12
for n in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29655+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=', d['n_gpus'], 'steps', d['steps'], 'ms_per_step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']/1e6,1), 'clocks', d['clocks'])"
done
