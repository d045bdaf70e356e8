#!/bin/bash
# stream kernels: tests, then A/B of the step with and without them
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "stream or k1_tiled or small_channels or dense" > gpurun_out/t_stream.log 2>&1; echo "stream tests rc=$?"; tail -15 gpurun_out/t_stream.log | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -x -q > gpurun_out/t_msdn.log 2>&1; echo "msdn rc=$?"; tail -3 gpurun_out/t_msdn.log | cut -c1-300
for cfg in "1 1" "0 0" "1 0" "0 1" "1 1" "0 0"; do
set -- $cfg
A3D_DENSE_STREAM=$1 A3D_K1_TILED=$2 A3D_AUTOTUNE_VERBOSE=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_s_$1$2.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('stream=$1 k1=$2', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1))
"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_s_$1$2.json
done
grep "kind 3\|kind 4" gpurun_out/bench_s_11.err | head
python - <<'P'
import json
for tag in ("11", "00"):
    d=json.load(open('gpurun_out/bench_ops_s_%s.json' % tag))
    print(tag, d['step_ms_graph'])
    for r in d['ops']:
        if 'dense_fwd' in r['op'] or 'dense_dgrad' in r['op'] or '55x74x1 ' in r['detail']: print('  %3d %-28s %-46s %8.4f' % (r['seq'], r['op'], r['detail'], r['ms']))
P
