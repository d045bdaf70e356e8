#!/bin/bash
mkdir -p gpurun_out
for c in "2 128" "0 128" "2 256"; do timeout 20 tools/probe/tmem_pair_probe $c >> gpurun_out/tmem_pair_probe.log 2>&1; echo " [rc=$?]" >> gpurun_out/tmem_pair_probe.log; done
cat gpurun_out/tmem_pair_probe.log
timeout 300 python tools/pair_rate_probe.py > gpurun_out/pair_rate_probe.log 2>&1; echo "rate probe rc=$?"; cat gpurun_out/pair_rate_probe.log | tail -70
