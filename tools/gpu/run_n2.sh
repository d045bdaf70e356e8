#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -5 gpurun_out/bench_n$N.err | cut -c1-300
python - <<P
import json
l=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print(l['n_gpus'], round(l['ms_per_step'],4), round(l['value']), round(l['e2e']['value']), l['config']['dp'])
P
