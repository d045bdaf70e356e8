#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "test_dense_fwd" 2>&1 | tail -2 | cut -c1-200
A3D_SWEEP_BATCHES=256,512 timeout 60 python tools/infer_sweep.py 2>&1 | tail -2 | cut -c1-200
