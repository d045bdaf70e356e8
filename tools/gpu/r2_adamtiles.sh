#!/bin/bash
mkdir -p gpurun_out
for t in 0 1 2 4 8 16; do for cfg in 230 231; do
  A3D_FUSED_ADAM_TILES=$t A3D_FUSED_ADAM_CFG=$cfg A3D_BENCH_U8=0 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t$t.json 2>/dev/null
  python - $t $cfg <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/bench_t{sys.argv[1]}.json").read().strip().splitlines()[-1])
ops=json.load(open("gpurun_out/bench_ops_latest.json"))["ops"]
fa=[round(o["ms"]*1e3,1) for o in ops if o["op"]=="a3d_dense_wgrad_adam"]
print("tiles/cta", sys.argv[1], "cfg", sys.argv[2], "ms/step", round(d["ms_per_step"],4), "fused alone us", fa)
PY
done; done
