#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$? in $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")})
print("e2e", d["e2e"]); print("e2e_u8", d["e2e_u8"]); print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"])
print("other", json.dumps(d.get("other_configs"), indent=1))
r=json.loads(open("gpurun_out/bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"])
PY
