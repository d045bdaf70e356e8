#!/bin/bash
for f in 0 1 0 1; do A3D_FUSE_DENSE_ADAM=$f timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fuse=$f', round(l['ms_per_step'], 4), round(l['value']), l['clocks'])
"; done
A3D_FUSE_DENSE_ADAM=1 timeout 600 python tools/ablate_step.py 30 2>&1 | grep -v "^a3d autotune" | cut -c1-200
