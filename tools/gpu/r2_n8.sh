#!/bin/bash
# 8 GPUs: timeline of the DP step, DP parity vs the oracle at batch 256, bench line
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 tools/step_timeline.py > gpurun_out/timeline_n$N.log 2>&1; echo "timeline rc=$?"
grep -A40 "^rank 0/" gpurun_out/timeline_n$N.log | head -45
timeout 300 $TR --master-port 29622 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("N", sys.argv[1], "ms/step", round(d["ms_per_step"],4), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
timeout 900 $TR --master-port 29623 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1; echo "dp_parity rc=$?"
grep "^{" gpurun_out/dp_parity_n$N.log
