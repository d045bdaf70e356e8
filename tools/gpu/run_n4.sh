#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
for ar in 0 1; do
A3D_DP_CONV_ALLREDUCE=$ar timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 60 --warmup 5 2>gpurun_out/bench_n${N}_ar$ar.err > gpurun_out/bench_n${N}_ar$ar.json
python -c "
import json
l = json.loads(open('gpurun_out/bench_n${N}_ar$ar.json').read().strip().splitlines()[-1])
print('conv_allreduce=$ar n=$N', round(l['ms_per_step'], 4), round(l['value']), 'e2e', round(l['e2e']['value']))
"
done
