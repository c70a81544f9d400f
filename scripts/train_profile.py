"""Per-kernel device time of one rgba_vae training step (torch.profiler / CUPTI; cheap, for iteration -- the judged
artifacts are the ncu launch lists under profiles/).  Usage: python scripts/train_profile.py [size] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import ragb_vae_b200 as R
from ragb_vae_b200.trainer import VaeTrainStep

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
vae = R.RgbaAutoencoder("qwen").to("cuda", torch.bfloat16)
step = VaeTrainStep(vae, loss_module=R.AlphaVaeLoss(reduce_mean=True))
x = torch.rand(B, 4, S, S, device="cuda")
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda")
for _ in range(3):
    step.step(x, noise)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step.step(x, noise)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device us {tot:.0f}  kernels {sum(e.count for e in rows)}")
for e in rows[:45]:
    print(f"{e.device_time_total:10.1f} us {e.count:5d}  {e.key[:110]}")
