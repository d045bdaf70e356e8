"""Per-kernel CUDA-event timings of one DCNF train step at batch 16 (768 patches): where the 18.9 ms go."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 480, 640, 3, generator=g).cuda()
depths = (torch.rand(B, 480, 640, 1, generator=g) * 0.95 + 0.05).cuda()
op = models.dcnf(images, depths, train=True)
pp = glorot_params(5, "dcnf")
pp["pairwise/pairwise_layers/dense/kernel"].abs_()
op.net.load_params(pp)
op.run()
rows = bench.per_op_profile(op, torch)
tot = sum(r["ms"] for r in rows)
print("sum of kernels %.3f ms" % tot)
for r in rows:
    fl = bench.conv_flops(r["detail"]) if r["op"].startswith("a3d_conv2d") else None
    r["tflops"] = (fl * 1e-9 / r["ms"]) if fl else None
    if r["ms"] > 0.02:
        print('%3d %-28s %-46s %8.3f %s' % (r["seq"], r["op"], r["detail"], r["ms"],
                                            ("%.0f TF/s" % r["tflops"]) if r["tflops"] else ""))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "dcnf_op_times.json")
json.dump({"batch": B, "sum_ms": tot, "rows": rows}, open(out, "w"), indent=1)
