"""Validation metrics of the ``rgba_vae`` stage (src/training/rgba_vae_stage.py:712-795) on the GPU.

The reference materialises two composites per background and runs ~10 elementwise kernels per
metric in the model dtype (bf16 PSNR is then quantised to 0.25 dB); ``validation_metrics`` reads
recon and target once and accumulates every background's squared error and the alpha MAE in fp32.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence, Union

import torch

from . import ops
from .rgba_vae import RgbaVAE, _ensure_alpha, background_rgb

NAMED_BACKGROUNDS = {"white": 1.0, "black": 0.0}


def resolve_background_spec(name: Union[str, float, Sequence[float]]):
    """rgba_vae_stage.py:787-795 -- 'white' / 'black' or a number / RGB triple."""
    if isinstance(name, str):
        key = name.lower()
        if key not in NAMED_BACKGROUNDS:
            raise ValueError(f"Unknown background spec '{name}'.")
        return NAMED_BACKGROUNDS[key]
    return name


def validation_metrics(recon: torch.Tensor, inputs: torch.Tensor,
                       backgrounds: Sequence[Union[str, float, Sequence[float]]] = ("white", "black")) -> Dict[str, torch.Tensor]:
    """Per-sample ``{"psnr_<bg>": (B,), ..., "alpha_mae": (B,)}`` in fp32; recon / inputs (B,4,H,W) in [0,1]."""
    specs = [background_rgb(resolve_background_spec(b)) for b in backgrounds]
    out = ops.composite_psnr(recon, inputs, specs)
    res = {f"psnr_{b}": out[:, i] for i, b in enumerate(backgrounds)}   # keys as the reference prints them (:770)
    res["alpha_mae"] = out[:, len(specs)]
    return res


def compute_psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """rgba_vae_stage.py:712-715 for already-composited (B,3,H,W) pairs: an opaque RGBA pair over any
    background is the pair itself, so the fused kernel is reused with alpha = 1."""
    if pred.shape != target.shape or pred.dim() != 4 or pred.shape[1] != 3:
        raise ValueError("compute_psnr expects two (B,3,H,W) tensors")
    one = torch.ones_like(pred[:, :1])
    return ops.composite_psnr(torch.cat([pred, one], 1), torch.cat([target, one], 1), [(0.0, 0.0, 0.0)])[:, 0]


@torch.no_grad()
def evaluate_rgba_vae(model: RgbaVAE, batches: Iterable[torch.Tensor],
                      backgrounds: Sequence[Union[str, float, Sequence[float]]] = ("white", "black"),
                      noises: Optional[Iterable[torch.Tensor]] = None) -> Dict[str, float]:
    """The loop of evaluate_rgba_vae (rgba_vae_stage.py:718-784): recon = model(inputs), composite
    over each background, PSNR + alpha MAE per sample, mean over all samples.  Metrics stay on the
    device until the end (the reference gathers + .cpu()s every batch)."""
    names = [f"psnr_{b}" for b in backgrounds] + ["alpha_mae"]
    sums, count = None, 0
    noise_it = iter(noises) if noises is not None else None
    for inputs in batches:   # inputs are used as they come, like the reference (no clamp)
        noise = next(noise_it) if noise_it is not None else None
        recon, _ = model(inputs, noise=noise)
        m = validation_metrics(recon, _ensure_alpha(inputs), backgrounds)
        vec = torch.stack([m[k].double().sum() for k in names])
        sums = vec if sums is None else sums + vec
        count += inputs.shape[0]
    if sums is None:
        return {}
    return dict(zip(names, (sums / count).cpu().tolist()))
