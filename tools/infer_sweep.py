"""BASELINE.json config 5: inference-only MSDN depth-map throughput sweep on one B200 (latency and images/s)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params

FWD_FLOP = 4.110e9
rows = []
p = glorot_params(1)
BATCHES = tuple(int(b) for b in os.environ.get("A3D_SWEEP_BATCHES", "1,2,4,8,16,32,64,128,256,512").split(","))
for bs in BATCHES:
    g = torch.Generator().manual_seed(bs)
    images = torch.rand(bs, 480, 640, 3, generator=g).cuda()
    depths = torch.zeros(bs, 55, 73, 1).cuda()
    op = models.msdn(images, depths, train=False)
    op.net.load_params(p)
    for _ in range(3):
        op.run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        op.net.forward()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    rows.append({"batch": bs, "latency_ms": ms, "images_per_s": bs / ms * 1e3, "tflops": FWD_FLOP * bs / ms / 1e9})
    print(rows[-1], flush=True)
    del op, gr
    torch.cuda.empty_cache()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "infer_sweep.json")
json.dump(rows, open(out, "w"), indent=1)
