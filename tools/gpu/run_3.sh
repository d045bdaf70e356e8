#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; grep -E "^FAILED|passed|failed|rror|timed out" gpurun_out/$name.log | cut -c1-240 | tail -12; }
run k_conv_fwd "conv_fwd"
run k_conv_bwd "dgrad_wgrad"
timeout 400 python -m pytest tests/test_gpu_msdn.py -m gpu -q -s 2>&1 | cut -c1-220 > gpurun_out/msdn6.log; grep -E "worst|passed|failed|FAILED|rror" gpurun_out/msdn6.log | tail -12
timeout 400 python -m pytest tests/test_gpu_dcnf.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -3 gpurun_out/bench4.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ops_latest.json'))
print(d['step_ms_graph'])
for r in d['ops']:
    if r['ms'] > 0.02: print('%3d %-28s %-46s %8.3f' % (r['seq'], r['op'], r['detail'], r['ms']))
"
