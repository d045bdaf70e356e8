#!/usr/bin/env python
"""Engine probe: tcgen05 GEMM rate for every combination of operand majors (K-major / MN-major).

    python tools/gemm_major_probe.py [M N K]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ann3depth_b200 import models  # noqa: E402


def main():
    M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 4096, 4096)
    ctx = models.get_context(0)
    dev = torch.device("cuda:0")
    A = (torch.rand(M, K, device=dev) - 0.5).bfloat16()
    B = (torch.rand(N, K, device=dev) - 0.5).bfloat16()
    At, Bt = A.t().contiguous(), B.t().contiguous()
    ref = A.float() @ B.float().t()
    for a_mn, b_mn, bn in ((False, False, 128), (False, False, 256), (True, False, 128), (False, True, 128),
                           (False, True, 256), (True, True, 128), (True, True, 256), (True, True, 64)):
        a = At if a_mn else A
        b = Bt if b_mn else B
        D = ctx.debug_tc_gemm(a, b, M, N, K, bn, 128, a_mn, b_mn, 1)
        torch.cuda.synchronize()
        err = float((D - ref).abs().max() / ref.abs().max())
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.debug_tc_gemm(a, b, M, N, K, bn, 128, a_mn, b_mn, 1)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps({"a_mn": a_mn, "b_mn": b_mn, "bn": bn, "ms": best, "tflops": 2.0 * M * N * K / best / 1e9,
                          "rel_err": err}), flush=True)


if __name__ == "__main__":
    main()
