#!/bin/bash
# 2 GPUs: new single-GPU tests first, then DP parity and the DP bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q -k "uint8 or hooks" > gpurun_out/t_msdn_new.log 2>&1; echo "new msdn tests rc=$?"; tail -3 gpurun_out/t_msdn_new.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/dp_parity.py 2>&1 | grep -v "^\*\|OMP\|NCCL version\|^$" | tail -4 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
l = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('n=2', round(l['ms_per_step'], 4), round(l['value']), 'e2e', round(l['e2e']['value']))
"
tail -2 gpurun_out/bench_n2.err | cut -c1-200
