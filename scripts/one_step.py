#!/usr/bin/env python
"""One RgbaVAE step (encode -> sample -> decode -> white-bg PSNR) for profiler captures.
    python scripts/one_step.py [arch] [batch] [size] [iters] [--loss]   (--loss: also the AlphaVAE reconstruction loss)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R
from ragb_vae_b200 import ops

with_loss = "--loss" in sys.argv
sys.argv = [a for a in sys.argv if a != "--loss"]
arch = sys.argv[1] if len(sys.argv) > 1 else "qwen"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
torch.manual_seed(0)
model = R.RgbaVAE(R.RgbaAutoencoder(arch).to("cuda", torch.bfloat16))
x = torch.rand(B, 4, S, S, device="cuda").bfloat16()
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda").bfloat16()
per_step = 0
for it in range(iters):
    if it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()   # `ncu --profile-from-start off` then sees exactly the last (warm) step
    l0 = ops.launch_count()
    recon, _ = model(x, noise=noise)
    m = ops.composite_psnr(recon, x, [(1.0, 1.0, 1.0)])
    if with_loss:
        R.AlphaVaeLoss(reduce_mean=True).reconstruction_loss(recon, x)   # same pass over the pair as the train step's loss
    per_step = ops.launch_count() - l0
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("psnr_white", m[:, 0].tolist(), "launches_total", ops.launch_count(), "launches_last_step", per_step)
