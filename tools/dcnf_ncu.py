"""One DCNF train step (batch 16) inside a cudaProfilerStart/Stop range, after two untimed steps (autotuning): the target
of `ncu --profile-from-start off` (tools/gpu/run_ncu_dcnf.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import models
from ann3depth_b200.init import glorot_params

B = 16
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 480, 640, 3, generator=g).cuda()
depths = (torch.rand(B, 480, 640, 1, generator=g) * 0.95 + 0.05).cuda()
op = models.dcnf(images, depths, train=True)
pp = glorot_params(5, "dcnf")
pp["pairwise/pairwise_layers/dense/kernel"].abs_()
op.net.load_params(pp)
for _ in range(2):
    op.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
op.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(op.net.loss))
