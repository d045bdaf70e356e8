#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msdn.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_all.log | cut -c1-300
for i in 1 2; do timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(l['ms_per_step'], 4), round(l['value']), l['clocks'], l['roofline']['conv_tensor_tflops'])
"; done
timeout 600 python tools/ablate_step.py 30 2>&1 | grep -v "^a3d autotune" | cut -c1-200
