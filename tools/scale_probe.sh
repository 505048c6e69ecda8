python tools/probe_step_gaps.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/probe_step_gaps.py 2>/dev/null
PROBE_SYNC=none python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 tools/probe_step_gaps.py 2>/dev/null
