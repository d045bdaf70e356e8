#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pool4 or resize_s2d or gather" 2>&1 | tail -3 | cut -c1-300
timeout 600 python tools/ablate_step.py 30 2>&1 | grep -v "^a3d autotune" | cut -c1-200
