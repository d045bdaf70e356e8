#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cta_pair or dgrad_act" > gpurun_out/t_pair.log 2>&1; echo "pair tests rc=$?"; tail -5 gpurun_out/t_pair.log
timeout 300 python tools/pair_rate_probe.py > gpurun_out/pair_rate_probe.log 2>&1; echo "rate rc=$?"
grep -E "variant.: (0|11|12)," gpurun_out/pair_rate_probe.log | grep -E "bn.: (128|256|192)"
