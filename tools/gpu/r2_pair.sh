#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/probe/tmem_pair_probe > gpurun_out/tmem_pair_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/tmem_pair_probe.log
A3D_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cta_pair and (12 or 7)" > gpurun_out/t_pair_deep.log 2>&1; echo "pair deep rc=$?"; tail -15 gpurun_out/t_pair_deep.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cta_pair or tile_variants" > gpurun_out/t_pair.log 2>&1; echo "pair all rc=$?"; tail -15 gpurun_out/t_pair.log
A3D_AUTOTUNE_VERBOSE=2 timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench pair rc=$?"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_pair.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_pair.json").read().strip().splitlines()[-1])
    print("pair", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["conv_tensor_tflops"])
except Exception as e: print("ERR", e)
PY
grep -- "-> candidate" gpurun_out/bench_pair.err | head -40
