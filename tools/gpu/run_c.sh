#!/bin/bash
mkdir -p gpurun_out
for h in 0 2; do
A3D_HALO=$h timeout 300 python tools/conv_sweep.py --reps 5 --out gpurun_out/sweep_halo$h.json > gpurun_out/sweep_halo$h.log 2>&1; echo "sweep halo=$h rc=$?"
done
python - <<'P'
import json
a={r['name']:r for r in json.load(open('gpurun_out/sweep_halo0.json'))}
b={r['name']:r for r in json.load(open('gpurun_out/sweep_halo2.json'))}
for n in a:
    ra, rb = a[n], b.get(n, {})
    s = '%-40s' % n[:40]
    for k in ('fwd_us','dgrad_us','wgrad_us'):
        if k in ra: s += ' %s %7.1f -> %7.1f' % (k[:-3], ra[k], rb.get(k, float('nan')))
    if 'error' in ra or 'error' in rb: s += ' ERR ' + str(ra.get('error', rb.get('error')))[:120]
    print(s)
P
