#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "resize" 2>&1 | tail -2 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn rc=$?"; tail -2 gpurun_out/t_msdn.log | cut -c1-300
for pr in 0 -1 0 -1; do A3D_MAIN_PRIORITY=$pr timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('prio=$pr', round(l['ms_per_step'], 4), round(l['value']), round(l['e2e']['value']))
"; done
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
for r in d['ops']:
    if 'resize' in r['op']: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
P
