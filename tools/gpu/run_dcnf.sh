#!/bin/bash
# DCNF first-layer rework: kernel tests of the overlapped view, the DCNF parity tests, per-op timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "view or patches or window" > gpurun_out/t_view.log 2>&1; echo "view tests rc=$?"
tail -15 gpurun_out/t_view.log
timeout 600 python -m pytest tests/test_gpu_dcnf.py tests/test_gpu_c_model.py -m gpu -q -x -k "dcnf" > gpurun_out/t_dcnf.log 2>&1; echo "dcnf tests rc=$?"
tail -15 gpurun_out/t_dcnf.log
timeout 250 python tools/dcnf_op_times.py 16 > gpurun_out/dcnf_op_times.log 2>&1; echo "op times rc=$?"
tail -45 gpurun_out/dcnf_op_times.log
