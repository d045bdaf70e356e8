#!/usr/bin/env python
"""Times single liba3d convolution / dense launches (fwd, dgrad, wgrad) on synthetic tensors.

Design experiments: which layout (channel padding, space-to-depth factor) the tcgen05 engine handles
best for a given layer.  Every sample is preceded by a 512 MB L2-flushing write and timed with CUDA
events on the launching stream; the minimum of `--reps` samples is printed.

    python tools/conv_sweep.py [--reps 5] [--only NAME_SUBSTR]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ann3depth_b200 import models, ops  # noqa: E402

# name, (N,H,W,C), K, R, S, stride, padding, f32 output, which passes
SHAPES = [
    # ---- current MSDN layers
    ("c0 11x12x4 s4 (virtual 16ch)", (32, 228, 304, 4), 96, 11, 12, 4, "valid", True, "fw"),
    ("c1 5x5 C96", (32, 27, 37, 96), 256, 5, 5, 1, "same", True, "fdw"),
    ("c2 3x3 C256", (32, 13, 18, 256), 384, 3, 3, 1, "same", False, "fdw"),
    ("c3 3x3 C384", (32, 13, 18, 384), 384, 3, 3, 1, "same", False, "fdw"),
    ("c4 3x3 s2 C384", (32, 13, 18, 384), 256, 3, 3, 2, "valid", False, "fdw"),
    ("f1 5x5x16 (s2d2)", (32, 114, 152, 16), 64, 5, 5, 1, "valid", True, "fw"),
    ("f2 5x5 C64", (32, 55, 74, 64), 64, 5, 5, 1, "same", False, "fdw"),
    # ---- candidates
    ("c0 as 3x3x64 over s2d4", (32, 57, 76, 64), 96, 3, 3, 1, "valid", True, "fw"),
    ("f1+pool as 3x3x64 -> 256 over s2d4", (32, 57, 76, 64), 256, 3, 3, 1, "valid", False, "fw"),
    ("c1 5x5 C128 (padded)", (32, 27, 37, 128), 256, 5, 5, 1, "same", True, "fdw"),
]


def flush(buf):
    buf.fill_(1.0)


def timed(fn, reps, buf):
    best = 1e9
    for _ in range(reps):
        flush(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3       # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ctx = models.get_context(0)
    dev = torch.device("cuda:0")
    buf = torch.empty(128 << 20, dtype=torch.float32, device=dev)
    rows = []
    for name, (N, H, W, Cc), K, R, S, stride, pad, f32, passes in SHAPES:
        if args.only and args.only not in name:
            continue
        d = ops.conv_desc(N, H, W, Cc, K, R, S, stride, pad)
        x = (torch.rand(N, H, W, Cc, device=dev) - 0.5).bfloat16()
        w = (torch.rand(K, R, S, Cc, device=dev) - 0.5).bfloat16()
        b = torch.zeros(K, device=dev)
        y = torch.empty(N, d.P, d.Q, K, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
        dy = (torch.rand(N, d.P, d.Q, K, device=dev) - 0.5).bfloat16()
        dx = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device=dev)
        dw = torch.empty(K, R, S, Cc, dtype=torch.float32, device=dev)
        db = torch.empty(K, dtype=torch.float32, device=dev)
        flops = 2.0 * N * d.P * d.Q * K * R * S * Cc
        rec = {"name": name, "gflop": flops / 1e9}
        try:
            if "f" in passes:
                fn = lambda: ctx.conv2d_fwd(d, x, w, b, relu=True, out=y)
                fn(); torch.cuda.synchronize()
                rec["fwd_us"] = timed(fn, args.reps, buf)
            if "d" in passes:
                fn = lambda: ctx.conv2d_dgrad(d, dy, w, out=dx)
                fn(); torch.cuda.synchronize()
                rec["dgrad_us"] = timed(fn, args.reps, buf)
            if "w" in passes:
                fn = lambda: ctx.conv2d_wgrad(d, x, dy, dw=dw, db=db)
                fn(); torch.cuda.synchronize()
                rec["wgrad_us"] = timed(fn, args.reps, buf)
        except Exception as e:          # keep sweeping: an unsupported candidate is a result too
            rec["error"] = str(e)[:200]
        for k in ("fwd", "dgrad", "wgrad"):
            if k + "_us" in rec:
                rec[k + "_tflops"] = flops / rec[k + "_us"] / 1e6
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    # pool-fused fine/first (conv + ReLU + 2x2 max-pool in one GEMM)
    if not args.only or "pool4" in args.only:
        d = ops.conv_desc(32, 57, 76, 64, 256, 3, 3, 1, "valid", ldy=64)
        x = (torch.rand(32, 57, 76, 64, device=dev) - 0.5).bfloat16()
        w = (torch.rand(256, 3, 3, 64, device=dev) - 0.5).bfloat16()
        b = torch.zeros(64, device=dev)
        y = torch.empty(32, 55, 74, 64, dtype=torch.bfloat16, device=dev)
        idx = torch.empty(32, 55, 74, 64, dtype=torch.uint8, device=dev)
        rec = {"name": "f1 pool4 fused 3x3x64 -> 4x64 (s2d4)"}
        fn = lambda: ctx.conv2d_pool4_fwd(d, x, w, b, relu=True, out=y, idx=idx)
        fn(); torch.cuda.synchronize()
        rec["fwd_us"] = timed(fn, args.reps, buf)
        fn = lambda: ctx.conv2d_pool4_fwd(d, x, w, b, relu=True, out=y, idx=None)
        rec["fwd_noidx_us"] = timed(fn, args.reps, buf)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    # dense layers (batch 32)
    for (M, Nn, K) in ((32, 4096, 12288), (32, 4070, 4096)):
        if args.only and "dense" not in args.only:
            continue
        x = (torch.rand(M, K, device=dev) - 0.5).bfloat16()
        w = (torch.rand(Nn, K, device=dev) - 0.5).bfloat16()
        b = torch.zeros(Nn, device=dev)
        ldy = (Nn + 63) // 64 * 64
        dy = (torch.rand(M, ldy, device=dev) - 0.5).bfloat16()
        y = torch.empty(M, Nn, dtype=torch.bfloat16, device=dev)
        dx = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
        dw = torch.empty(Nn, K, dtype=torch.float32, device=dev)
        db = torch.empty(Nn, dtype=torch.float32, device=dev)
        rec = {"name": f"dense {M}x{Nn}x{K}"}
        for key, fn in (("fwd_us", lambda: ctx.dense_fwd(x, w, b, out=y)),
                        ("dgrad_us", lambda: ctx.dense_dgrad(dy, w, out=dx)),
                        ("wgrad_us", lambda: ctx.dense_wgrad(x, dy, dw=dw, db=db, N=Nn))):
            fn(); torch.cuda.synchronize()
            rec[key] = timed(fn, args.reps, buf)
        rec["weight_MB_bf16"] = Nn * K * 2 / 1e6
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
