#!/usr/bin/env python
"""Config c5: Flux-arch RGBA decode of one 2048x2048 image (z 1x16x256x256) as inference_rgba_flux.py does:
decode(z/scale + shift).sample -> (y+1)/2 -> clamp.  Prints time and peak memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R

torch.manual_seed(0)
vae = R.RgbaAutoencoder("flux").to("cuda", torch.bfloat16)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
z = (torch.randn(1, 16, S // 8, S // 8, generator=torch.Generator().manual_seed(3)) / 0.3611 + 0.1159).cuda().bfloat16()
zn = ((z.float() - 0.1159) * 0.3611).bfloat16()
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    img = vae._decode_image(zn, out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0), z_scale=1.0 / 0.3611, z_shift=0.1159)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run {i}: {dt*1e3:.1f} ms  -> {tuple(img.shape)} finite={bool(torch.isfinite(img.float()).all())} "
          f"range=[{float(img.min()):.3f},{float(img.max()):.3f}] peak_mem={torch.cuda.max_memory_allocated()/2**30:.1f} GiB "
          f"({48.496/dt:.0f} TFLOP/s algorithmic)" if S == 2048 else f"run {i}: {dt*1e3:.1f} ms")
