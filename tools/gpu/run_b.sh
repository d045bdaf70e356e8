#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_major_probe.py 4096 4096 4096 2>&1 | tee gpurun_out/major_probe.log | cut -c1-200
