#!/usr/bin/env python
"""GroupNorm(32) + SiLU forward / backward kernels alone at a Flux full-resolution shape -- development aid.
    python scripts/experiments/gn_bwd_time.py [C] [side] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ragb_vae_b200 import ops
from ragb_vae_b200 import training as T

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, S, S, C, device="cuda", generator=g).bfloat16()
dy = torch.randn(B, S, S, C, device="cuda", generator=g).bfloat16()
add = torch.randn(B, S, S, C, device="cuda", generator=g).bfloat16()
gamma = torch.nn.Parameter(torch.rand(C, device="cuda") + 0.5)
beta = torch.nn.Parameter(torch.rand(C, device="cuda") - 0.5)
gw, gb = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
y, stats = ops.groupnorm_silu(x, gw, gb, 32, 1e-6, True, return_stats=True)
dg = torch.zeros(C, device="cuda")
db = torch.zeros(C, device="cuda")
el = x.numel() * 2


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ms = timed(lambda: ops.groupnorm_silu(x, gw, gb, 32, 1e-6, True))
print(f"forward  (stats + apply, 3E)        {ms:.3f} ms  {3 * el / ms / 1e6:7.1f} GB/s")
ms = timed(lambda: T.groupnorm_silu_backward(x, stats, gamma, beta, dy, 32, 1e-6, True, dgamma_out=dg, dbeta_out=db))
print(f"backward (reduce + apply, 5E)       {ms:.3f} ms  {5 * el / ms / 1e6:7.1f} GB/s")
ms = timed(lambda: T.groupnorm_silu_backward(x, stats, gamma, beta, dy, 32, 1e-6, True, dgamma_out=dg, dbeta_out=db, add=add))
print(f"backward + skip branch (6E)         {ms:.3f} ms  {6 * el / ms / 1e6:7.1f} GB/s")
