#!/usr/bin/env python
"""Engine rate of the GEMM tile variants on plain K-major GEMMs (CUDA events, 20 launches): one-CTA tiles (variant 0/1/2)
against the CTA-pair kernel (11 short ring, 12 deep ring).  Not a bench value -- a probe for DESIGN.md section 4.1."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import models

ctx = models.get_context(0)
out = []
for (M, N, K) in ((4096, 4096, 4096), (32768, 256, 3200), (7488, 384, 3456), (130240, 64, 1600)):
    A = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    ref = None
    for bn in (64, 128, 192, 256):
        if bn > N and bn != 64:
            continue
        for variant in (0, 1, 2, 11, 12):
            try:
                D = ctx.debug_tc_gemm(A, B, M, N, K, bn, 128, variant=variant)
            except Exception as e:
                continue
            torch.cuda.synchronize()
            if ref is None:
                ref = D.clone()
            err = float((D - ref).abs().max() / ref.abs().max())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ctx.debug_tc_gemm(A, B, M, N, K, bn, 128, variant=variant)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            row = dict(M=M, N=N, K=K, bn=bn, variant=variant, us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6, 1), err=err)
            print(row, flush=True)
            out.append(row)
json.dump(out, open("gpurun_out/pair_rate_probe.json", "w"), indent=1)
