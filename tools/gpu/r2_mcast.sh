#!/bin/bash
# round 2, call 1: does the cluster/multicast kernel work on hardware, and does it win?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
A3D_TEST_MCAST=1 A3D_MCAST=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "multicast" > gpurun_out/t_mcast.log 2>&1; echo "mcast tests rc=$?"; tail -5 gpurun_out/t_mcast.log
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "bench base rc=$?"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_base.json
A3D_MCAST=1 A3D_AUTOTUNE_VERBOSE=2 timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mcast.json 2> gpurun_out/bench_mcast.err; echo "bench mcast rc=$?"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_mcast.json
python - <<'PY'
import json
for n in ("base","mcast"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["conv_tensor_tflops"])
    except Exception as e: print(n, "ERR", e)
PY
