#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn rc=$?"; tail -3 gpurun_out/t_msdn.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -2 gpurun_out/t_kernels.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -3 gpurun_out/bench_i.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
print('step_ms_graph', d['step_ms_graph'], 'sum', sum(r['ms'] for r in d['ops']))
for r in d['ops']:
    if r['ms'] > 0.015: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
l=json.loads(open('gpurun_out/bench_i.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline'])
P
