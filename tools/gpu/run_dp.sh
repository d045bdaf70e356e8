#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29631 tools/h2d_ceiling.py > gpurun_out/h2d_n$N.log 2>&1; echo "h2d rc=$?"; grep "^{" gpurun_out/h2d_n$N.log
timeout 300 $TR --master-port 29632 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
A3D_NUMA_BIND=0 timeout 300 $TR --master-port 29633 bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/bench_n${N}_unbound.json 2> gpurun_out/bench_n${N}_unbound.err; echo "bench unbound rc=$?"
python - $N <<'PY'
import json,sys
for tag in ("", "_unbound"):
    d=json.loads(open(f"gpurun_out/bench_n{sys.argv[1]}{tag}.json").read().strip().splitlines()[-1])
    print("N", sys.argv[1], tag, "ms/step", round(d["ms_per_step"],4), "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
PY
