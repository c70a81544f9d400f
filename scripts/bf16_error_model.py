#!/usr/bin/env python
"""CPU model of where bf16 storage error enters the GPU pipeline (development aid, not a test).

Runs the fp32 oracle with hooks that round tensors to bf16 at the points where the CUDA path stores
bf16 (norm outputs = MMA A operands, conv weights, block outputs = residual stream, q/k/v/P/O in the
attention), with switches to keep selected tensors in fp32, and prints the relative error of the
decoded image / moments against the unmodified oracle."""
import sys, os, itertools
import torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vae_oracle as O

def r(t): return t.bfloat16().float()

def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm())

def instrument(vae, *, weights=True, norm_out=True, conv_out=True, stream=True, stream_fp32_min_c=None, attn=True):
    """stream_fp32_min_c: keep the residual stream / conv outputs in fp32 where channels >= this value."""
    hooks = []
    if weights:
        for m in vae.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
                m.weight.data = r(m.weight.data)
    def keep32(t):
        return stream_fp32_min_c is not None and t.shape[1] >= stream_fp32_min_c
    def hook_round(mod, inp, out):
        return out if keep32(out) else r(out)
    def hook_round_always(mod, inp, out):
        return r(out)
    for name, m in vae.named_modules():
        if isinstance(m, (nn.GroupNorm, O.QwenRMSNorm)) and norm_out:
            # norm output itself is not stored; silu(norm) is -> round after silu: emulate by rounding here (close enough)
            hooks.append(m.register_forward_hook(hook_round_always))
        elif isinstance(m, (O.FluxResnetBlock2D, O.QwenResidualBlock, O.FluxAttention, O.QwenAttentionBlock)) and stream:
            hooks.append(m.register_forward_hook(hook_round))
        elif isinstance(m, (nn.Conv2d, nn.Conv3d)) and conv_out:
            last = name.split(".")[-1]
            if last == "conv2" or last == "proj":   # fused with the residual add in fp32
                continue
            hooks.append(m.register_forward_hook(hook_round))
        elif isinstance(m, nn.Linear) and attn:
            if name.endswith("to_out.0"): continue
            hooks.append(m.register_forward_hook(hook_round_always))
    return hooks

def main():
    torch.set_num_threads(8)
    arch = sys.argv[1] if len(sys.argv) > 1 else "qwen"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ref = O.build_oracle(arch, 0)
    x = O.synthetic_rgba(1, size, size, seed=1)
    noise = torch.randn(1, 16, size//8, size//8, generator=torch.Generator().manual_seed(2))
    recon, post, z = O.rgba_vae_forward(ref, x, noise)
    dec = ref.decode(z).sample
    variants = {
        "all bf16 (GPU pipeline model)": dict(),
        "weights only": dict(norm_out=False, conv_out=False, stream=False, attn=False),
        "norm_out only": dict(weights=False, conv_out=False, stream=False, attn=False),
        "conv_out only": dict(weights=False, norm_out=False, stream=False, attn=False),
        "stream only": dict(weights=False, norm_out=False, conv_out=False, attn=False),
        "all, fp32 stream+conv_out where C>=384": dict(stream_fp32_min_c=384),
        "all, fp32 stream+conv_out where C>=192": dict(stream_fp32_min_c=192),
        "all, fp32 stream+conv_out everywhere": dict(stream_fp32_min_c=1),
    }
    for name, kw in variants.items():
        m = O.build_oracle(arch, 0)
        instrument(m, **kw)
        mom = m.encode_moments(O.to_vae_range(x))
        d = m.decode(r(z)).sample
        print(f"{arch} {size}  {name:45s} moments {rel(mom, post.parameters):.4f}  decoded {rel(d, dec):.4f}", flush=True)

if __name__ == "__main__":
    main()
