#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wgrad or prepared" > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -8 gpurun_out/t_kernels.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn rc=$?"; tail -3 gpurun_out/t_msdn.log | cut -c1-300
for i in 1 2 3; do
A3D_AUTOTUNE_VERBOSE=2 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_u_$i.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('run $i', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1), round(l['e2e']['value']))
"
done
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_u.json
grep "kind 1 \[" gpurun_out/bench_u_1.err | cut -c1-160
