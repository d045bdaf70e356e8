#!/bin/bash
# 2 GPUs: DP parity against the oracle at batch 64, then the bench line at N=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dp_parity.py > gpurun_out/dp_parity_n2.log 2>&1; echo "dp_parity rc=$?"
tail -5 gpurun_out/dp_parity_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_n2.json
