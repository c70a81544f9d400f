#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (run once, in the build container).

For each arch the oracle (oracle/vae_oracle.py) is random-initialised under seed 0 and run on
config c1 of BASELINE.json (1x4x256x256 fp32, supplied noise).  For ``flux`` the same weights
are ALSO loaded, through a key remap, into the independent BFL auto-encoder that ships with
torchtitan in this image; its moments / decoded sample are stored next to the oracle's so
that tests can pin the oracle against a second implementation without torchtitan installed.

The reference itself (jaejung-dev/ragb-vae) cannot generate vectors: it imports ``diffusers``
at module top (src/models/rgba_vae.py:17, losses.py:7) and diffusers is not installable here.
"""
from __future__ import annotations

import json
import os
import sys

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vae_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def titan_state_dict(sd):
    """diffusers AutoencoderKL keys -> BFL ``AutoEncoder`` keys."""
    out = {}
    for k, v in sd.items():
        side, rest = k.split(".", 1)
        r = rest
        r = r.replace("conv_norm_out", "norm_out")
        r = r.replace("mid_block.resnets.0", "mid.block_1").replace("mid_block.resnets.1", "mid.block_2")
        r = r.replace("mid_block.attentions.0", "mid.attn_1")
        r = r.replace("group_norm", "norm").replace("to_q", "q").replace("to_k", "k").replace("to_v", "v")
        r = r.replace("to_out.0", "proj_out")
        r = r.replace("conv_shortcut", "nin_shortcut")
        if r.startswith("down_blocks."):
            _, i, kind, j, tail = r.split(".", 4)
            r = f"down.{i}.block.{j}.{tail}" if kind == "resnets" else f"down.{i}.downsample.{tail}"
        if r.startswith("up_blocks."):
            _, i, kind, j, tail = r.split(".", 4)
            i = 3 - int(i)
            r = f"up.{i}.block.{j}.{tail}" if kind == "resnets" else f"up.{i}.upsample.{tail}"
        if "attn_1" in r and r.endswith("weight") and v.dim() == 2:
            v = v[:, :, None, None]
        out[f"{side}.{r}"] = v.clone()
    return out


def run(arch: str):
    vae = O.build_oracle(arch, seed=0)
    x = O.synthetic_rgba(1, 256, 256, seed=1)
    noise = torch.randn(1, 16, 32, 32, generator=torch.Generator().manual_seed(2))
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    recon, post, z = O.rgba_vae_forward(vae, x, noise)
    dec_raw = vae.decode(z).sample
    tgt = O.to_vae_range(x)
    metrics = O.validation_metrics(recon, x)
    tensors = {
        "moments": post.parameters.contiguous(),
        "z": z.contiguous(),
        "decoded": dec_raw.contiguous(),
        "recon": recon.contiguous(),
        "recon_loss_sum": O.reconstruction_loss(dec_raw, tgt, reduce_mean=False).reshape(1),
        "recon_loss_mean": O.reconstruction_loss(dec_raw, tgt, reduce_mean=True).reshape(1),
        "naive_mse_mean": O.reconstruction_loss(dec_raw, tgt, reduce_mean=True, use_naive_mse=True).reshape(1),
        "kl": post.kl().contiguous(),
        "psnr_white": metrics[1.0].contiguous(),
        "psnr_black": metrics[0.0].contiguous(),
        "alpha_mae": metrics["alpha_mae"].contiguous(),
        "x_checksum": x.double().sum().float().reshape(1),
        "noise_checksum": noise.double().sum().float().reshape(1),
        "weight_checksum": sum(p.double().abs().sum() for p in vae.parameters()).float().reshape(1),
    }
    info = {"arch": arch, "params": sum(p.numel() for p in vae.parameters())}
    if arch == "flux":
        from torchtitan.experiments.flux.model.autoencoder import AutoEncoder, AutoEncoderParams

        ae = AutoEncoder(AutoEncoderParams(in_channels=4, out_ch=4)).eval().requires_grad_(False)
        missing = ae.load_state_dict(titan_state_dict(vae.state_dict()), strict=True)
        t_mom = ae.encoder(O.to_vae_range(x))
        t_dec = ae.decoder(z)
        tensors["titan_moments"] = t_mom.contiguous()
        tensors["titan_decoded"] = t_dec.contiguous()
        info["titan_vs_oracle_moments_maxabs"] = float((t_mom - post.parameters).abs().max())
        info["titan_vs_oracle_decoded_maxabs"] = float((t_dec - dec_raw).abs().max())
        info["titan_load"] = str(missing)
    save_file(tensors, os.path.join(GOLD, f"{arch}_c1_256.safetensors"))
    info.update({k: [float(t) for t in v.flatten()[:4]] for k, v in tensors.items() if v.numel() <= 4})
    return info


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"torch": torch.__version__, "runs": [run(a) for a in ("qwen", "flux")]}
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))
