"""DiagonalGaussianDistribution on the GPU (diffusers ``models/autoencoders/vae.py``).

Reference uses: ``vae.encode(x).latent_dist.sample()`` (src/models/rgba_vae.py:278,
src/training/rgba_vae_stage.py:451, src/models/flux_kontext_textalpha.py:331), ``.kl(other)``
(rgba_vae.py:314, losses.py:114), ``.parameters`` and the constructor
``DiagonalGaussianDistribution(tensor)`` (rgba_vae_stage.py:696-700).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        if parameters.dim() != 4 or parameters.shape[1] % 2:
            raise ValueError(f"moments must be (B, 2*Z, H, W), got {tuple(parameters.shape)}")
        self.parameters = parameters
        self.deterministic = deterministic
        self._cache = {}

    # chunk / clamp / exp views are only materialised when somebody asks for them; the hot
    # path (sample / kl) reads `parameters` directly inside one fused kernel.
    def _get(self, name):
        if name not in self._cache:
            mean, logvar = torch.chunk(self.parameters, 2, dim=1)
            logvar = torch.clamp(logvar, -30.0, 20.0)
            self._cache["mean"] = mean
            self._cache["logvar"] = logvar
            if self.deterministic:
                self._cache["std"] = self._cache["var"] = torch.zeros_like(mean)
            else:
                self._cache["std"] = torch.exp(0.5 * logvar)
                self._cache["var"] = torch.exp(logvar)
        return self._cache[name]

    mean = property(lambda self: self._get("mean"))
    logvar = property(lambda self: self._get("logvar"))
    std = property(lambda self: self._get("std"))
    var = property(lambda self: self._get("var"))

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None,
               shift: float = 0.0, scale: float = 1.0) -> torch.Tensor:
        """mean + std * eps.  ``noise`` (our extension) supplies eps so that results are reproducible
        against the oracle; without it eps is drawn like diffusers' ``randn_tensor``.  ``shift`` /
        ``scale`` fuse the Flux latent normalisation ``(z - shift) * scale``."""
        p = self.parameters
        n, c2, h, w = p.shape
        if self.deterministic:
            return ((self.mean - shift) * scale).clone()
        if noise is None:
            noise = torch.randn((n, c2 // 2, h, w), generator=generator, device=p.device, dtype=p.dtype)
        z, _ = ops.reparam(p, noise, want_kl=False, z_shift=shift, z_scale=scale)
        return z

    def kl(self, other: "Optional[DiagonalGaussianDistribution]" = None) -> torch.Tensor:
        if self.deterministic:
            return torch.zeros(1, device=self.parameters.device)
        if other is None:
            _, kl = ops.reparam(self.parameters, None, want_kl=True)
            return kl
        # KL to another posterior (reference-KL term, rgba_vae_stage.py:489-508): one fused pass (rv_kl_ref)
        if other.deterministic:
            raise ValueError("kl(other): the reference posterior is deterministic (zero variance)")
        return ops.kl_to_reference(self.parameters, other.parameters)[0]

    def mode(self) -> torch.Tensor:
        return self.mean
