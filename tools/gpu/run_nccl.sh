#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for ch in default 4 8 16; do
if [ "$ch" = "default" ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nchannels=$ch', l['n_gpus'], round(l['ms_per_step'], 4), round(l['value']))
"
done
