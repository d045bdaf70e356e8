#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rows or rank_blocks or adam" 2>&1 | tail -2 | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/dp_parity.py 2>&1 | grep -v "^\*\|OMP\|NCCL version\|^$" | tail -4 | cut -c1-300
for ar in 0 1; do
A3D_DP_CONV_ALLREDUCE=$ar timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('conv_allreduce=$ar', l['n_gpus'], round(l['ms_per_step'], 4), round(l['value']))
"
done
