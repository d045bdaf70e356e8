"""Product-side parameter initialiser (tf.layers defaults: glorot-uniform kernels, zero biases) in the
reference's TF variable layouts.  Lives outside `oracle/` so that bench.py's GPU arm and the train
driver never import the oracle."""
import math
from collections import OrderedDict

import torch

from ann3depth_b200.params import dcnf_specs, msdn_specs


def glorot_params(seed=1, model="msdn", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    specs = msdn_specs() if model == "msdn" else dcnf_specs()
    # draw in the TF variable-creation order (sorted by name keeps it deterministic)
    out = OrderedDict()
    for s in sorted(specs, key=lambda s: s.name):
        if s.kind == "bias":
            out[s.name] = torch.zeros(s.tf_shape, dtype=dtype)
            continue
        shape = s.tf_shape
        rf = 1
        for d in shape[:-2]:
            rf *= d
        lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
        out[s.name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
    return out
