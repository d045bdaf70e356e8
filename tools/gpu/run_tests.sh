#!/bin/bash
# full GPU suite (what the driver runs at round end) + timing of the slowest tests
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/t_gpu_all.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/t_gpu_all.log
