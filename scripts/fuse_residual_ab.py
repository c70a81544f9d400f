#!/usr/bin/env python
"""A/B of RgbaAutoencoder.fuse_norm_residual (conv2 + residual also writes the next block's act(norm1(.))) on config c2:
CUDA-graph replay, L2 flushed between steps.   python scripts/fuse_residual_ab.py [arch] [batch] [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R

arch = sys.argv[1] if len(sys.argv) > 1 else "qwen"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
torch.manual_seed(0)
vae = R.RgbaAutoencoder(arch).to("cuda", torch.bfloat16)
model = R.RgbaVAE(vae)
x = torch.rand(B, 4, S, S, device="cuda").bfloat16()
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for rounds in range(2):
    for flag in (False, True):
        vae.fuse_norm_residual = flag
        model.reset_graphs()
        for _ in range(3):
            recon, _, met = model.forward_graphed(x, noise)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            flush.zero_()
            recon, _, met = model.forward_graphed(x, noise)
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = recon.float().clone()
        err = float((recon.float() - ref).norm() / ref.norm())
        print(f"fuse_norm_residual={flag}: {e0.elapsed_time(e1) / 8:.2f} ms/step  psnr {float(met[:, 0].mean()):.4f}  rel diff vs first {err:.2e}")
