#!/bin/bash
mkdir -p gpurun_out
for cfg in 230 231 220 221 210 211 130 131 121 111; do
  A3D_FUSED_ADAM_CFG=$cfg A3D_BENCH_U8=0 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$cfg.json 2>/dev/null
  python - $cfg <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_cfg{sys.argv[1]}.json").read().strip().splitlines()[-1])
ops=json.load(open("gpurun_out/bench_ops_latest.json"))["ops"]
fa=[round(o["ms"]*1e3,1) for o in ops if o["op"]=="a3d_dense_wgrad_adam"]
print("cfg", sys.argv[1], "ms/step", round(d["ms_per_step"],4), "fused alone us", fa)
PY
done
