#!/usr/bin/env python
"""RMS-norm + SiLU kernel alone at the shapes of the c2 step -- development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ragb_vae_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (b, s, c) in ((8, 128, 384), (8, 256, 384), (8, 512, 192), (8, 1024, 96)):
    x = torch.randn(b, s, s, c, device="cuda").bfloat16()
    g = torch.rand(c, device="cuda") + 0.5
    y = torch.empty_like(x)
    ops.rmsnorm_silu(x, g, True, out=y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.rmsnorm_silu(x, g, True, out=y)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"rmsnorm+silu ({b},{s},{s},{c}): {ms * 1e3:7.1f} us  {2 * x.numel() * 2 / ms / 1e6:7.1f} GB/s")
