#!/bin/bash
mkdir -p gpurun_out
A3D_FUSED_ADAM=simt timeout 300 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q -k "fused or general_adam" 2>&1 | tail -2 | cut -c1-300
A3D_FUSED_ADAM=simt timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -3 gpurun_out/bench_j.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
print('step_ms_graph', d['step_ms_graph'], 'sum', sum(r['ms'] for r in d['ops']))
for r in d['ops']:
    if 'adam' in r['op']: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
P
