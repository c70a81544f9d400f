import sys, os, torch
sys.path.insert(0, "/root/repo")
import ragb_vae_b200 as R
from ragb_vae_b200 import ops
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n, t = 4, 16384
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(n * t, d, device="cuda", generator=g).bfloat16()
k = torch.randn(n * t, d, device="cuda", generator=g).bfloat16()
vt = torch.randn(n, d, t, device="cuda", generator=g).bfloat16()
for _ in range(2):
    o = ops.attention(q, k, vt, n, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    o = ops.attention(q, k, vt, n, t)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"d={d} n={n} t={t}: {ms:.3f} ms  {4.0*n*t*t*d/ms/1e9:.1f} TFLOP/s")
