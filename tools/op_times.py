"""Per-kernel CUDA-event timings of one sequential phase-1 step (same table bench.py writes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params
images, depths = bench.synthetic_batch(0, torch)
op = models.msdn(images.cuda(), depths.cuda(), train=True, overlap=False)
op.net.load_params(glorot_params(seed=1))
op.run(use_graph=False)
rows = bench.per_op_profile(op, torch)
tot = sum(r["ms"] for r in rows)
print("sum of kernels %.3f ms" % tot)
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for r in rows:
    if r["ms"] > 0.03 and flt in r["op"]:
        print('%3d %-28s %-46s %8.3f' % (r["seq"], r["op"], r["detail"], r["ms"]))
