"""The steps either side of the VAE on the GPU: the train step's triplet detail augmentation and
posterior split (src/training/rgba_vae_stage.py:606-625,690-700), the Flux latent patchify of
src/models/flux_kontext_textalpha.py:330-349, and the uint8 RGBA <-> tensor conversions of
inference_rgba_flux.py:15-26, and the batch assembly + RandomBackgroundBlend augmentation of
rgba_vae_stage.py:85-130,575-603.  Same names and errors as the reference functions; one kernel each."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check
from .ops import _dt, _need_cuda, _ptr, _stream
from .posterior import DiagonalGaussianDistribution


def build_detail_augmented_triplet(target: torch.Tensor) -> torch.Tensor:
    """(B,4,H,W) in [-1,1] -> (3B,4,H,W): original, composited on black, composited on white (alpha = 1)."""
    if target.dim() != 4 or target.shape[1] < 4:
        raise ValueError("detail augmentation expects RGBA tensors.")
    _need_cuda(target)
    target = target.contiguous()
    b, _, h, w = target.shape
    out = torch.empty((3 * b, 4, h, w), dtype=target.dtype, device=target.device)
    check(_lib.load().rv_triplet_augment(_ptr(target), _ptr(out), b, h * w, _dt(target), _stream(target)), "rv_triplet_augment")
    return out


class RandomBackgroundBlend:
    """rgba_vae_stage.py:85-130 for a whole batch on the device: each sample is, with probability ``prob``, composited
    over a random opaque colour drawn uniformly from ``color_range`` (alpha becomes 1).  ``__call__`` takes the (B,4,H,W)
    RGBA batch in [0,1] and returns ``(batch, background_augmented)`` -- the bool mask the reference stores per sample
    under ``"background_augmented"``.  ``mask`` / ``colors`` can be supplied for reproducibility."""

    def __init__(self, prob: float = 0.1, color_range: Tuple[float, float] = (0.2, 0.9)) -> None:
        if color_range[0] >= color_range[1]:
            raise ValueError("color_range lower bound must be < upper bound.")
        self.prob = prob
        self.color_range = tuple(float(v) for v in color_range)

    def __call__(self, batch: torch.Tensor, generator=None, mask: Optional[torch.Tensor] = None,
                 colors: Optional[torch.Tensor] = None):
        if batch.dim() != 4 or batch.shape[1] != 4:
            raise ValueError("background blend expects a (B,4,H,W) RGBA batch.")
        _need_cuda(batch)
        batch = batch.contiguous()
        b, _, h, w = batch.shape
        dev = batch.device
        if mask is None:
            mask = torch.rand(b, generator=generator, device=dev) < self.prob
        if colors is None:
            lo, hi = self.color_range
            colors = torch.rand(b, 3, generator=generator, device=dev) * (hi - lo) + lo
        mask_u8 = mask.to(device=dev, dtype=torch.uint8).contiguous()
        colors = colors.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty_like(batch)
        check(_lib.load().rv_background_blend(_ptr(batch), _ptr(colors), _ptr(mask_u8), _ptr(out), b, h * w, _dt(batch),
                                              _stream(batch)), "rv_background_blend")
        return out, mask_u8.bool()


def build_training_batch(batch: dict, device, *, background_sample_prob: float = 0.0, generator=None,
                         background_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgba_vae_stage.py:575-603: cat(component, composite) (or composite alone) on ``device``, plus the background frames
    a Bernoulli(background_sample_prob) mask selects (``background_mask`` supplies the draw, for reproducibility).  Same
    ValueErrors as the reference."""
    tensors = []
    if "component" in batch and "composite" in batch:
        tensors.extend([batch["component"], batch["composite"]])
    elif "composite" in batch:
        tensors.append(batch["composite"])
    else:
        raise ValueError("Batch must contain 'composite' tensor for training.")
    inputs = torch.cat([t.to(device, non_blocking=True) for t in tensors], dim=0)
    if (background_sample_prob > 0.0 or background_mask is not None) and "background" in batch:
        background = batch["background"].to(device, non_blocking=True)
        if background.dim() == 3:
            background = background.unsqueeze(0)
        if background.shape[1] != 4:
            raise ValueError("Background tensor is expected to have 4 channels (RGBA).")
        if background_mask is not None:
            mask = background_mask.to(device=background.device, dtype=torch.bool)
        else:
            mask = torch.rand(background.shape[0], device=device, generator=generator) < background_sample_prob
        if mask.any():
            inputs = torch.cat([inputs, background[mask]], dim=0)
    return inputs


def split_triplet_distribution(posterior: DiagonalGaussianDistribution) -> Tuple[DiagonalGaussianDistribution, ...]:
    """rgba_vae_stage.py:690-700: chunk the moments in three along the batch and rebuild three posteriors."""
    chunks = torch.chunk(posterior.parameters, 3, dim=0)
    if len(chunks) != 3 or posterior.parameters.shape[0] % 3:
        raise ValueError("Posterior batch dimension must be divisible by 3 for triplet splits.")
    return tuple(DiagonalGaussianDistribution(c) for c in chunks)


def pack_latents(latents: torch.Tensor, shift: float = 0.0, scale: float = 1.0) -> torch.Tensor:
    """FluxPipeline._pack_latents: (B,C,h,w) -> (B,(h/2)(w/2),4C), optionally (z - shift) * scale fused."""
    _need_cuda(latents)
    latents = latents.contiguous()
    b, c, h, w = latents.shape
    out = torch.empty((b, (h // 2) * (w // 2), 4 * c), dtype=latents.dtype, device=latents.device)
    check(_lib.load().rv_pack_latents(_ptr(latents), _ptr(out), b, c, h, w, _dt(latents), shift, scale, 0, _stream(latents)),
          "rv_pack_latents")
    return out


def unpack_latents(tokens: torch.Tensor, height: int, width: int, vae_scale_factor: int = 8, shift: float = 0.0,
                   scale: float = 1.0) -> torch.Tensor:
    """FluxPipeline._unpack_latents: (B,T,4C) -> (B,C,h,w) for a height x width pixel image, optionally
    z * scale + shift fused (the ``latents / scaling_factor + shift_factor`` in front of decode)."""
    _need_cuda(tokens)
    tokens = tokens.contiguous()
    b, t, f = tokens.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    if t != (h // 2) * (w // 2) or f % 4:
        raise ValueError(f"token tensor {tuple(tokens.shape)} does not match a {height}x{width} image")
    c = f // 4
    out = torch.empty((b, c, h, w), dtype=tokens.dtype, device=tokens.device)
    check(_lib.load().rv_pack_latents(_ptr(tokens), _ptr(out), b, c, h, w, _dt(tokens), shift, scale, 1, _stream(tokens)),
          "rv_pack_latents")
    return out


def rgba_u8_to_tensor(img_u8: torch.Tensor, dtype: torch.dtype = torch.bfloat16, vae_range: bool = False) -> torch.Tensor:
    """(B,H,W,4) or (H,W,4) uint8 -> (B,4,H,W) in [0,1] (or [-1,1] with ``vae_range``): load_rgba + _to_vae_range."""
    _need_cuda(img_u8)
    if img_u8.dtype != torch.uint8 or img_u8.shape[-1] != 4:
        raise ValueError("expected a uint8 RGBA image tensor (..., H, W, 4)")
    if img_u8.dim() == 3:
        img_u8 = img_u8.unsqueeze(0)
    img_u8 = img_u8.contiguous()
    b, h, w, _ = img_u8.shape
    out = torch.empty((b, 4, h, w), dtype=dtype, device=img_u8.device)
    scale, shift = (2.0, -1.0) if vae_range else (1.0, 0.0)
    check(_lib.load().rv_rgba_u8_to_nchw(_ptr(img_u8), _ptr(out), b, h * w, _dt(out), scale, shift, _stream(img_u8)),
          "rv_rgba_u8_to_nchw")
    return out


def tensor_to_rgba_u8(x: torch.Tensor) -> torch.Tensor:
    """(B,4,H,W) in [0,1] -> (B,H,W,4) uint8, clamp + *255 + truncate like save_rgba."""
    _need_cuda(x)
    if x.dim() != 4 or x.shape[1] != 4:
        raise ValueError("expected a (B,4,H,W) RGBA tensor")
    x = x.contiguous()
    b, _, h, w = x.shape
    out = torch.empty((b, h, w, 4), dtype=torch.uint8, device=x.device)
    check(_lib.load().rv_nchw_to_rgba_u8(_ptr(x), _ptr(out), b, h * w, _dt(x), _stream(x)), "rv_nchw_to_rgba_u8")
    return out
