#!/bin/bash
# 2-GPU check: default DP conv path (bf16 reduce-scatter + sharded Adam) against the replicated f32 all-reduce path
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29641 bench.py --gpus 2 --steps 200 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
A3D_DP_CONV_ALLREDUCE=1 timeout 300 $TR --master-port 29642 bench.py --gpus 2 --steps 200 --warmup 3 > gpurun_out/bench_n2_allreduce.json 2> gpurun_out/bench_n2_allreduce.err; echo "bench allreduce rc=$?"
python - <<'PY'
import json
for tag in ("", "_allreduce"):
    d=json.loads(open(f"gpurun_out/bench_n2{tag}.json").read().strip().splitlines()[-1])
    print("N 2", tag, "ms/step", round(d["ms_per_step"],4), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
