"""AlphaVaeLoss on the GPU -- mirror of src/models/losses.py (reconstruction / KL terms).

``reconstruction_loss`` is one fused pass over pred and target (``rv_recon_loss``) instead of the
reference's ~12 elementwise kernels + ``.mean()``.  LPIPS (a third-party VGG, losses.py:85-107) is
outside the hot path and is not provided.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .posterior import DiagonalGaussianDistribution

DEFAULT_EB = (-0.0357, -0.0811, -0.1797)   # losses.py:34-37
DEFAULT_EB2 = (0.3163, 0.3060, 0.3634)


class AlphaVaeLoss(nn.Module):
    def __init__(self, *, reduce_mean: bool = False, use_naive_mse: bool = False, use_lpips: bool = False,
                 custom_eb: Optional[Sequence[float]] = None, custom_eb2: Optional[Sequence[float]] = None) -> None:
        super().__init__()
        custom_eb = DEFAULT_EB if custom_eb is None else custom_eb
        custom_eb2 = DEFAULT_EB2 if custom_eb2 is None else custom_eb2
        if len(custom_eb) != 3 or len(custom_eb2) != 3:
            raise ValueError("custom_eb/custom_eb2 must each provide three channel weights.")
        if use_lpips:
            raise ImportError("LPIPS is a third-party perceptual network outside the RGBA-VAE hot path; "
                              "ragb_vae_b200 does not provide it (set lpips_scale to 0).")
        self.reduce_mean = reduce_mean
        self.use_naive_mse = use_naive_mse
        self.use_lpips = False
        self.register_buffer("eb", torch.tensor(custom_eb, dtype=torch.float32).view(1, 3, 1, 1), persistent=False)
        self.register_buffer("eb2", torch.tensor(custom_eb2, dtype=torch.float32).view(1, 3, 1, 1), persistent=False)
        self._eb, self._eb2 = tuple(float(v) for v in custom_eb), tuple(float(v) for v in custom_eb2)

    def reconstruction_loss(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """``pred`` / ``target`` in [-1, 1], channels RGBA (losses.py:67-83).  Returns an fp32 scalar."""
        per_sample = ops.recon_loss_per_sample(pred, target, self._eb, self._eb2, self.use_naive_mse)
        if self.reduce_mean:  # mean over the (B,3,H,W) loss map -- (B,4,H,W) for the naive MSE
            ch = 4 if self.use_naive_mse else 3
            return per_sample.sum() / float(pred.shape[0] * ch * pred.shape[2] * pred.shape[3])
        return per_sample.mean()  # per-sample sum, then batch mean (losses.py:117-123)

    def kl_loss(self, posterior: DiagonalGaussianDistribution,
                reference: Optional[DiagonalGaussianDistribution] = None) -> torch.Tensor:
        return self._reduce(posterior.kl(reference))

    def _reduce(self, value: torch.Tensor) -> torch.Tensor:
        if value.ndim == 0:
            return value
        if self.reduce_mean:
            return value.mean()
        return value.view(value.shape[0], -1).sum(dim=1).mean()
