#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "resize" > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -8 gpurun_out/t_kernels.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn rc=$?"; tail -3 gpurun_out/t_msdn.log | cut -c1-300
for pl in 1 0 1; do
A3D_RESIZE_PIPELINED=$pl timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_v_$pl.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pipelined=$pl', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1), 'e2e', round(l['e2e']['value']), 'e2e_u8', l['e2e_u8'] and round(l['e2e_u8']['value']), l['e2e_u8'] and l['e2e_u8']['last_loss'], l['e2e']['last_loss'])
"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
for r in d['ops']:
    if 'resize' in r['op']: print('   %-28s %8.4f' % (r['op'], r['ms']))
P
done
tail -3 gpurun_out/bench_v_1.err | cut -c1-300
