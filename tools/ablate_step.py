#!/usr/bin/env python
"""Marginal cost of each group of launches inside the captured, multi-stream phase-1 step.

For every group the launches are replaced by no-ops, the CUDA graph is re-captured and the step is timed:
step(all) - step(without group) is what the group costs ON THE CRITICAL PATH (as opposed to its stand-alone
duration in profiles/bench_ops_*.json).  Results are numerically meaningless; timing only.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ann3depth_b200 import models  # noqa: E402
from ann3depth_b200.init import glorot_params  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda:0")
    images, depths = bench.synthetic_batch(0, torch)
    op = models.msdn(images.to(dev), depths.to(dev), train=True)
    net, ctx = op.net, op.net.ctx
    net.load_params(glorot_params(seed=1))
    fine_descs = (net.d_f2, net.d_f3)

    def skip_fine_convs(orig):
        def f(d, *a, **k):
            if d is net.d_f2 or d is net.d_f3:
                return None
            return orig(d, *a, **k)
        return f

    noop = lambda *a, **k: None
    groups = {
        "none": {},
        "adam": {"adam_tf": noop},
        "fused dense wgrad+adam (both layers)": {"dense_wgrad_adam": noop},
        "dense_wgrad": {"dense_wgrad": noop},
        "dense_fwd+dgrad": {"dense_fwd": noop, "dense_dgrad": noop},
        "conv_wgrad": {"conv2d_wgrad": noop},
        "conv_dgrad": {"conv2d_dgrad": noop},
        "fine (pool4+f2+f3)": {"conv2d_pool4_fwd": noop, "conv2d_fwd": skip_fine_convs},
        "resize": {"resize_bilinear_tf1_s2d": noop, "resize_bilinear_tf1": noop},
        "pools": {"maxpool2x2_fwd_f32": noop, "maxpool2x2_idx_bwd": noop},
        "adam+dense_wgrad": {"adam_tf": noop, "dense_wgrad": noop},
        "all conv bwd": {"conv2d_wgrad": noop, "conv2d_dgrad": noop},
    }
    base = None
    out = []
    for name, patch in groups.items():
        saved = {}
        for meth, repl in patch.items():
            saved[meth] = getattr(ctx, meth)
            setattr(ctx, meth, repl(saved[meth]) if repl is skip_fine_convs else repl)
        net._graphs.clear()
        for _ in range(3):
            op.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            op.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        for meth, fn in saved.items():
            setattr(ctx, meth, fn)
        if base is None:
            base = ms
        rec = {"without": name, "ms_per_step": round(ms, 4), "marginal_ms": round(base - ms, 4)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ablate_step.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
