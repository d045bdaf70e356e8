#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_parity.py 2>&1 | grep -v "^\*\|OMP\|NCCL version\|^$" | tail -4 | cut -c1-300
for sp in 1 0 1; do
A3D_DP_CONV_SPLIT=$sp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 2>gpurun_out/bench_n2_$sp.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split=$sp n=2', round(l['ms_per_step'], 4), round(l['value']), 'e2e', round(l['e2e']['value']))
"
done
