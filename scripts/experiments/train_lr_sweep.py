import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ragb_vae_b200 as R
from ragb_vae_b200.trainer import VaeTrainStep
from oracle import vae_oracle as O
for lr in (1e-5, 3e-5):
    oracle = O.build_oracle("qwen", seed=0)
    vae = R.RgbaAutoencoder("qwen"); vae.load_state_dict(oracle.state_dict()); vae = vae.to("cuda", torch.bfloat16)
    step = VaeTrainStep(vae, lr=lr, kl_scale=1e-6, loss_module=R.AlphaVaeLoss(reduce_mean=True))
    x = O.synthetic_rgba(4, 64, 64, seed=51, structured=True).cuda()
    noise = torch.randn(4, 16, 8, 8, generator=torch.Generator().manual_seed(52)).cuda()
    ls = [float(step.step_graphed(x, noise)["train/recon"]) for _ in range(80)]
    print(lr, [round(v, 4) for v in ls[::5]])
