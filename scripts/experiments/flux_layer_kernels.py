#!/usr/bin/env python
"""One Flux 256 -> 256 @512^2 layer through the kernels added in round 2's second session: the conv with the GroupNorm
statistics epilogue (+ finishing kernel), GroupNorm apply with those statistics, GroupNorm backward, the weight-gradient
kernel (N = 192 multi-tap MMAs) -- the command behind profiles/r02b_ncu_flux_layer_selected.csv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ragb_vae_b200 import _lib, ops
from ragb_vae_b200 import training as T

n, s, c = 4, 512, 256
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, s, s, c, device="cuda", generator=g).bfloat16()
dy = torch.randn(n, s, s, c, device="cuda", generator=g).bfloat16()
w = torch.randn(c, c, 3, 3, device="cuda", generator=g) / (3 * c ** 0.5)
bias = torch.randn(c, device="cuda", generator=g)
gamma = torch.nn.Parameter(torch.rand(c, device="cuda") + 0.5)
beta = torch.nn.Parameter(torch.rand(c, device="cuda") - 0.5)
wp = ops.pack_conv_weights_tc(w)
desc = ops.make_desc(n, s, s, c, c, 3, 1, False, x_dtype=_lib.RV_BF16, y_dtype=_lib.RV_BF16)
y = torch.empty(n, s, s, c, dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    stats = ops.conv2d_tc_gnstats(desc, x, wp, wp.shape[1], bias, None, y, 32)
    a = ops.groupnorm_silu(y, gamma.detach().float(), beta.detach().float(), 32, 1e-6, True, stats=stats)
    dx, _, _ = T.groupnorm_silu_backward(y, stats, gamma, beta, dy, 32, 1e-6, True)
    dw, db = T.conv_wgrad(a, dy, 3)
torch.cuda.synchronize()
print("ok", float(a.float().abs().mean()), float(dw.abs().mean()))
