#!/usr/bin/env python
"""Times rv_gemm_rowstat (the two score-matrix GEMMs of the attention backward) at the c4 shape -- development aid.
    python scripts/experiments/rowstat_time.py [tokens] [d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ragb_vae_b200 import ops

t = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(t, d, device="cuda", generator=g).bfloat16()
k = torch.randn(t, d, device="cuda", generator=g).bfloat16()
v = torch.randn(t, d, device="cuda", generator=g).bfloat16()
do = torch.randn(t, d, device="cuda", generator=g).bfloat16()
scale = d ** -0.5
o, lse = ops.attention(q, k, v.t().contiguous().view(1, d, t), 1, t, return_lse=True)
delta = ops.rowdot(do, o, scale)
p = torch.empty(t, t, device="cuda", dtype=torch.bfloat16)
ds = torch.empty_like(p)


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


fl = 2.0 * t * t * d
ms1 = timed(lambda: ops.gemm_rowstat(q, k, lse, 1, scale * 1.4426950408889634, out=p))
ms2 = timed(lambda: ops.gemm_rowstat(do, v, delta, 2, scale, mul_in=p, out=ds))
print(f"P  = exp2(QK^T - lse)      {t}x{t}x{d}: {ms1:.3f} ms  {fl / ms1 / 1e9:7.1f} TFLOP/s  {t * t * 2 / ms1 / 1e6:7.1f} GB/s written")
print(f"dS = P * (dO V^T - delta)  {t}x{t}x{d}: {ms2:.3f} ms  {fl / ms2 / 1e9:7.1f} TFLOP/s  {t * t * 4 / ms2 / 1e6:7.1f} GB/s moved")
