#!/bin/bash
# first GPU validation: each group in its own process so a faulting kernel cannot poison the others
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -25 gpurun_out/$name.log; }
run elementwise "resize or maxpool or relu_bwd or silog or adam or cast or crf or pairwise"
run tc_kmajor "tc_gemm_kmajor"
run tc_mn "tc_gemm_mn_major"
run conv_simt "conv_fwd and simt or small_channels"
run conv_tc "conv_fwd and tc or concat"
run conv_bwd "dgrad_wgrad"
run dense "dense"
