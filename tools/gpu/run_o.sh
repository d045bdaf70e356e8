#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "adam or dense" > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_k.log | cut -c1-300
for g in 1 0; do
A3D_DP_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_g$g.json 2> gpurun_out/bench_n2_g$g.err; echo "gather=$g rc=$?"
tail -3 gpurun_out/bench_n2_g$g.err | grep -v "^\*\|OMP\|NCCL version" | cut -c1-300
python - <<P
import json
l=json.loads(open('gpurun_out/bench_n2_g$g.json').read().strip().splitlines()[-1])
print('gather=$g', l['n_gpus'], round(l['ms_per_step'],4), round(l['value']), round(l['e2e']['value']), l['e2e'].get('last_loss'))
P
done
