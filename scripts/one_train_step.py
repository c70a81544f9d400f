#!/usr/bin/env python
"""Eager rgba_vae training steps (VaeTrainStep.step) for profiler captures.
    python scripts/one_train_step.py [size] [batch] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R
from ragb_vae_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
vae = R.RgbaAutoencoder("qwen").to("cuda", torch.bfloat16)
step = R.VaeTrainStep(vae, loss_module=R.AlphaVaeLoss(reduce_mean=True))
x = torch.rand(B, 4, S, S, device="cuda")
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda")
for _ in range(iters):
    l0 = ops.launch_count()
    m = step.step(x, noise)
    per = ops.launch_count() - l0
torch.cuda.synchronize()
print("loss", float(m["train/loss"]), "librgbavae launches per step", per)
