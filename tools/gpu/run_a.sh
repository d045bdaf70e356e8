#!/bin/bash
# GPU call A: validate the TMA epilogue, sweep layer shapes with/without it, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -8 | cut -c1-300
A3D_EPI_TMA=0 timeout 300 python tools/conv_sweep.py --reps 5 --out gpurun_out/sweep_epi0.json > gpurun_out/sweep_epi0.log 2>&1; echo "sweep0 rc=$?"
A3D_EPI_TMA=1 timeout 300 python tools/conv_sweep.py --reps 5 --out gpurun_out/sweep_epi1.json > gpurun_out/sweep_epi1.log 2>&1; echo "sweep1 rc=$?"
python - <<'P'
import json
a={r['name']:r for r in json.load(open('gpurun_out/sweep_epi0.json'))}
b={r['name']:r for r in json.load(open('gpurun_out/sweep_epi1.json'))}
for n in a:
    ra, rb = a[n], b.get(n, {})
    s = '%-40s' % n[:40]
    for k in ('fwd_us','dgrad_us','wgrad_us'):
        if k in ra: s += ' %s %7.1f -> %7.1f' % (k[:-3], ra[k], rb.get(k, float('nan')))
    if 'error' in ra or 'error' in rb: s += ' ERR ' + str(ra.get('error', rb.get('error')))[:120]
    print(s)
P
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
print('step_ms_graph', d['step_ms_graph'], 'sum', sum(r['ms'] for r in d['ops']))
for r in d['ops']:
    if r['ms'] > 0.02: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
l=json.loads(open('gpurun_out/bench_a.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline'])
P
