#!/usr/bin/env python
"""RMS-norm + SiLU backward kernel alone at the c4 shapes -- development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ragb_vae_b200 import training as T

for (b, s, c) in ((4, 1024, 96), (4, 512, 192), (4, 256, 384)):
    x = torch.randn(b, s, s, c, device="cuda").bfloat16()
    dy = torch.randn(b, s, s, c, device="cuda").bfloat16()
    add = torch.randn(b, s, s, c, device="cuda").bfloat16()
    gamma = torch.nn.Parameter(torch.rand(c, 1, 1, device="cuda") + 0.5)
    dg = torch.zeros(c, device="cuda")
    for use_add in (False, True):
        fn = lambda: T.rmsnorm_silu_backward(x, gamma, dy, True, dgamma_out=dg, add=add if use_add else None)
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        n = (4 if use_add else 3) * x.numel() * 2
        print(f"rmsnorm+silu backward ({b},{s},{s},{c}) add={use_add}: {ms * 1e3:7.1f} us  {n / ms / 1e6:7.1f} GB/s")
