#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pool or relu or dgrad or dense_fwd or dense_bwd or resize or space" > gpurun_out/t_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -4 gpurun_out/t_kernels.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_msdn.py tests/test_gpu_dcnf.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn+dcnf rc=$?"; tail -3 gpurun_out/t_msdn.log | cut -c1-300
for i in 1 2; do
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_y_$i.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('run $i', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1), 'e2e', round(l['e2e']['value']), 'e2e_u8', l['e2e_u8'] and round(l['e2e_u8']['value']))
"
done
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
for r in d['ops']:
    if 'pool' in r['op'] or 'relu' in r['op'] or 'dgrad' in r['op'] or 'dense_fwd' in r['op']: print('   %-28s %-44s %8.4f' % (r['op'], r['detail'], r['ms']))
P
