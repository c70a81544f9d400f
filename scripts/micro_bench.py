#!/usr/bin/env python
"""Stand-alone timing of the HBM-bound kernels at config-c2 sizes (B=8, 1024^2, bf16): CUDA events, 20 iterations after
5 warm-ups, each call replayed from a CUDA graph (no host gaps), a 256 MiB L2 flush between iterations.  Prints achieved GB/s (algorithmic bytes) and the fraction of the
measured HBM peak in MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ragb_vae_b200 import ops, plumbing

peak = 6549.8
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, nbytes, name):
    """fn is captured into a CUDA graph (no host gaps between its launches); each replay is preceded by an L2 flush."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(5):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gbs = nbytes / 1e9 / (ms / 1e3)
    print(f"{name:34s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/peak:5.2f} of measured HBM peak ({nbytes/2**20:.0f} MiB algorithmic)")


B, H, W = 8, 1024, 1024
x = torch.rand(B, 4, H, W, device="cuda").bfloat16()
y = torch.rand(B, 4, H, W, device="cuda").bfloat16()
timeit(lambda: ops.composite_psnr(x, y, [(1.0, 1.0, 1.0)]), 2 * x.numel() * 2, "composite+PSNR (white) + alpha MAE")
timeit(lambda: ops.composite_psnr(x, y, [(1.0, 1.0, 1.0), (0.0, 0.0, 0.0)]), 2 * x.numel() * 2, "composite+PSNR (white+black)")
timeit(lambda: ops.recon_loss_per_sample(x, y, (-0.0357, -0.0811, -0.1797), (0.3163, 0.3060, 0.3634)), 2 * x.numel() * 2, "AlphaVAE recon loss")
# the same two reductions at 4x the batch: 128 MiB takes ~40 us, mostly ramp-up / tail / the finishing launch; 512 MiB shows the
# streaming rate of the kernel itself
xb = torch.rand(4 * B, 4, H, W, device="cuda").bfloat16()
yb = torch.rand(4 * B, 4, H, W, device="cuda").bfloat16()
timeit(lambda: ops.composite_psnr(xb, yb, [(1.0, 1.0, 1.0)]), 2 * xb.numel() * 2, "composite+PSNR (white), B = 32")
timeit(lambda: ops.recon_loss_per_sample(xb, yb, (-0.0357, -0.0811, -0.1797), (0.3163, 0.3060, 0.3634)), 2 * xb.numel() * 2, "AlphaVAE recon loss, B = 32")
del xb, yb
mom = torch.randn(B, 32, H // 8, W // 8, device="cuda").bfloat16()
eps = torch.randn(B, 16, H // 8, W // 8, device="cuda").bfloat16()
timeit(lambda: ops.reparam(mom, eps), (mom.numel() + 2 * eps.numel()) * 2, "reparameterize")
for c, hw in ((96, 1024), (192, 512), (384, 256)):
    a = torch.randn(B, hw, hw, c, device="cuda").bfloat16()
    g = torch.ones(c, device="cuda")
    out = torch.empty_like(a)
    timeit(lambda: ops.rmsnorm_silu(a, g, True, out), 2 * a.numel() * 2, f"RMS-norm+SiLU C={c} @{hw}^2")
    del a, out
a = torch.randn(4, 1024, 1024, 128, device="cuda").bfloat16()
g, b_ = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
out = torch.empty_like(a)
timeit(lambda: ops.groupnorm_silu(a, g, b_, 32, 1e-6, True, out), 3 * a.numel() * 2, "GroupNorm(32)+SiLU C=128 @1024^2 (B=4)")
t = (torch.rand(4, 4, H, W, device="cuda") * 2 - 1).bfloat16()
timeit(lambda: plumbing.build_detail_augmented_triplet(t), 4 * t.numel() * 2, "triplet augmentation (B=4)")
