import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ragb_vae_b200 as R
from ragb_vae_b200 import ops
torch.manual_seed(0)
for flag in (False, True):
    vae = R.RgbaAutoencoder("qwen").to("cuda", torch.bfloat16)
    vae.fuse_norm_residual = flag
    model = R.RgbaVAE(vae)
    x = torch.rand(8, 4, 1024, 1024, device="cuda").bfloat16()
    noise = torch.randn(8, 16, 128, 128, device="cuda").bfloat16()
    for _ in range(3):
        model.forward_graphed(x, noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        model.forward_graphed(x, noise)
    e1.record(); torch.cuda.synchronize()
    print("fuse_norm_residual", flag, e0.elapsed_time(e1) / 10, "ms/step")
    del model, vae
