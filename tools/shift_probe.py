"""Hardware probe: row-shifted UMMA descriptor starts inside a SWIZZLE_128B tile (see tc_shift_test.cu)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann3depth_b200 import _lib as L, ops

ctx = ops.Context(0)
g = torch.Generator().manual_seed(0)
N = 64
for a_mn in (0, 1):
    A = (torch.rand(72, 128, generator=g) - .5) if a_mn else (torch.rand(136, 64, generator=g) - .5)
    B = torch.rand(N, 64, generator=g) - .5
    Ab, Bb = A.to(torch.bfloat16).cuda(), B.to(torch.bfloat16).cuda()
    for shift in range(0, 9):
        ref = (Ab[shift:shift + 64].float().t() if a_mn else Ab[shift:shift + 128].float()) @ Bb.float().t()
        for ubo in (0, 1):
            D = torch.zeros(128, N, device="cuda")
            L.check(ctx.lib.a3d_debug_tc_shift(ctx.h, C.c_void_p(Ab.data_ptr()), C.c_void_p(Bb.data_ptr()),
                                               C.c_void_p(D.data_ptr()), N, shift, ubo, a_mn, None), "shift")
            torch.cuda.synchronize()
            err = float((D - ref).abs().max() / ref.abs().max())
            print(f"a_mn={a_mn} shift={shift} base_offset={'set' if ubo else '0  '} rel_err={err:.3e} {'OK' if err < 2e-3 else 'WRONG'}",
                  flush=True)
