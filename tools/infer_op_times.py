"""Per-kernel CUDA-event timings of one MSDN inference pass (BASELINE.json config 5) at a given batch size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator().manual_seed(bs)
im = torch.rand(bs, 480, 640, 3, generator=g).cuda()
op = models.msdn(im, torch.zeros(bs, 55, 73, 1, device="cuda"), train=False)
op.net.load_params(glorot_params(seed=1))
op.run()
rows = bench.per_op_profile(op, torch)
tot = sum(r["ms"] for r in rows)
print("batch %d: sum of kernels %.3f ms = %.0f img/s" % (bs, tot, bs / tot * 1e3))
for r in rows:
    fl = bench.conv_flops(r["detail"]) if r["op"].startswith("a3d_conv2d") else None
    r["tflops"] = (fl * 1e-9 / r["ms"]) if fl else None
    if r["ms"] > 0.01:
        print('%3d %-28s %-46s %8.3f %s' % (r["seq"], r["op"], r["detail"], r["ms"],
                                            ("%.0f TF/s" % r["tflops"]) if r["tflops"] else ""))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "infer_op_times_bs%d.json" % bs)
json.dump({"batch": bs, "sum_ms": tot, "rows": rows}, open(out, "w"), indent=1)
