#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -s > gpurun_out/t_parity_full.log 2>&1; echo "parity_full rc=$?"
tail -60 gpurun_out/t_parity_full.log
timeout 900 python -m pytest tests/test_gpu_driver.py -x -q > gpurun_out/t_driver.log 2>&1; echo "driver rc=$?"
tail -30 gpurun_out/t_driver.log
