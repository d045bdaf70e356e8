#!/bin/bash
for lib in "" ann3depth_b200/liba3d_r72.so ann3depth_b200/liba3d_r64.so "" ann3depth_b200/liba3d_r72.so; do
A3D_LIB=${lib:+$PWD/$lib} timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('lib=${lib:-default}', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1))
"
done
