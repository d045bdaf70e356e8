#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q 2>&1 | tail -8 | cut -c1-300
timeout 200 python tools/conv_sweep.py --reps 5 --only pool4 2>&1 | tail -2 | cut -c1-300
for cfg in "1 0" "1 3" "1 4" "1 6" "2 0" "2 3" "2 4" "2 6" "3 0" "3 4"; do set -- $cfg; echo "wgrad nblk=$1 splits=$2"; A3D_WGRAD_NBLK=$1 A3D_WGRAD_SPLITS=$2 timeout 200 python tools/conv_sweep.py --reps 5 --only " C" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('  %-30s wgrad %7.1f' % (r['name'][:30], r.get('wgrad_us', -1)), r.get('error', '')[:100])
"; done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -3 gpurun_out/bench_d.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
print('step_ms_graph', d['step_ms_graph'], 'sum', sum(r['ms'] for r in d['ops']))
for r in d['ops']:
    if r['ms'] > 0.015: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
l=json.loads(open('gpurun_out/bench_d.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline'])
P
