"""Host-side mirror of the reference's RGBA wrapper (src/models/rgba_vae.py) over RgbaAutoencoder.

Same names, argument meaning and error behaviour as the reference functions cited per symbol; the
arithmetic is librgbavae's.  ``RgbaVAE.forward`` folds the reference's three extra elementwise
passes (``_ensure_alpha`` / ``_to_vae_range`` before encode, ``_from_vae_range`` + clamp after
decode; rgba_vae.py:274-281) into the first conv's loader and the last conv's epilogue.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence, Union

import torch
import torch.nn as nn

from . import ops
from .autoencoder import CONFIG_NAME, WEIGHTS_NAME, RgbaAutoencoder

BackgroundSpec = Union[float, Sequence[float], torch.Tensor]


def _ensure_alpha(x: torch.Tensor) -> torch.Tensor:
    """rgba_vae.py:25-29 -- append alpha = 1 to a 3-channel batch."""
    if x.shape[1] == 4:
        return x
    if x.shape[1] != 3:
        raise ValueError(f"expected a 3- or 4-channel image batch, got {x.shape[1]} channels")
    return torch.cat([x, torch.ones_like(x[:, :1])], dim=1)


def _to_vae_range(x: torch.Tensor) -> torch.Tensor:  # rgba_vae.py:32-33
    return x * 2.0 - 1.0


def _from_vae_range(x: torch.Tensor) -> torch.Tensor:  # rgba_vae.py:36-37
    return (x + 1.0) * 0.5


def background_rgb(background: BackgroundSpec) -> Sequence[float]:
    """Scalar or RGB-triple background -> (r, g, b).  Mirrors the scalar / sequence branches of
    ``_normalize_background`` (rgba_vae.py:40-72); tensor backgrounds go through
    ``composite_over_background``."""
    if isinstance(background, torch.Tensor):
        raise TypeError("tensor backgrounds are only supported by composite_over_background")
    if isinstance(background, (list, tuple)):
        if len(background) != 3:
            raise ValueError("Background color sequence must contain exactly three values.")
        return tuple(float(v) for v in background)
    return (float(background),) * 3


def _normalize_background(background: BackgroundSpec, reference: torch.Tensor) -> torch.Tensor:
    """rgba_vae.py:40-72 (same ValueErrors); returns a broadcastable (B|1, 3, H|1, W|1) tensor."""
    b, _, h, w = reference.shape
    if isinstance(background, torch.Tensor):
        bg = background.to(device=reference.device, dtype=reference.dtype)
        if bg.dim() == 3:
            bg = bg.unsqueeze(0)
        if bg.dim() != 4:
            raise ValueError(f"Background tensor must have 3 or 4 dimensions, got {bg.dim()}")
        if bg.shape[0] == 1 and b > 1:
            bg = bg.expand(b, -1, -1, -1)
        if bg.shape[1] == 1:
            bg = bg.repeat(1, 3, 1, 1)
        if bg.shape[2] != h or bg.shape[3] != w:
            raise ValueError("Background tensor spatial size must match the RGBA tensor.")
        return bg
    rgb = background_rgb(background)
    return torch.tensor(rgb, device=reference.device, dtype=reference.dtype).view(1, 3, 1, 1)


def composite_over_background(rgba: torch.Tensor, background: BackgroundSpec) -> torch.Tensor:
    """rgba_vae.py:75-84 -- rgb*a + bg*(1-a) as a materialised (B,3,H,W) tensor.  Only the
    visualisation helpers need the image; validation uses the fused ``composite_psnr`` kernel, which
    never materialises it."""
    rgba = _ensure_alpha(rgba)
    rgb, alpha = rgba[:, :3], rgba[:, 3:4]
    return rgb * alpha + _normalize_background(background, rgb) * (1.0 - alpha)


def composite_over_white(rgba: torch.Tensor) -> torch.Tensor:  # rgba_vae.py:87-88
    return composite_over_background(rgba, 1.0)


def composite_over_black(rgba: torch.Tensor) -> torch.Tensor:  # rgba_vae.py:91-92
    return composite_over_background(rgba, 0.0)


def adapt_vae_to_rgba(vae, alpha_bias_init: float = 0.0) -> None:
    """rgba_vae.py:95-123 -- widen ``encoder.conv_in`` / ``decoder.conv_out`` from 3 to 4 channels:
    the new alpha input column and alpha output row are zero, ``bias[3] = alpha_bias_init``.
    Rank-agnostic (works on the 5-D causal-conv3d kernels of the Qwen arch)."""
    conv_in = vae.encoder.conv_in
    if conv_in.in_channels != 4:
        w = conv_in.weight.data
        nw = torch.zeros(w.size(0), 4, *w.shape[2:], dtype=w.dtype, device=w.device)
        nw[:, :3] = w
        conv_in.in_channels = 4
        conv_in.weight = nn.Parameter(nw, requires_grad=conv_in.weight.requires_grad)
    conv_out = vae.decoder.conv_out
    if conv_out.out_channels != 4:
        w = conv_out.weight.data
        nw = torch.zeros(4, w.size(1), *w.shape[2:], dtype=w.dtype, device=w.device)
        nw[:3] = w
        nb = torch.zeros(4, dtype=w.dtype, device=w.device)
        nb[:3] = conv_out.bias.data
        nb[3] = alpha_bias_init
        conv_out.out_channels = 4
        conv_out.weight = nn.Parameter(nw, requires_grad=conv_out.weight.requires_grad)
        conv_out.bias = nn.Parameter(nb, requires_grad=conv_out.bias.requires_grad)
    vae.config.in_channels = 4
    vae.config.out_channels = 4


def _maybe_restore_rgba_convs(vae, model_name_or_path: str, subfolder: Optional[str]) -> bool:
    """rgba_vae.py:143-191 -- a checkpoint saved after widening still says ``in_channels: 3`` in its
    config while its tensors are 4-channel; re-read those tensors straight from the safetensors file.
    Returns True if anything was restored; raises RuntimeError on NaN/Inf like the reference."""
    from safetensors import safe_open

    root = os.path.join(model_name_or_path, subfolder) if subfolder else model_name_or_path
    path = os.path.join(root, WEIGHTS_NAME)
    if not os.path.isfile(path):
        return False
    restored = False
    targets = {"encoder.conv_in.weight": (vae.encoder.conv_in, "weight"),
               "decoder.conv_out.weight": (vae.decoder.conv_out, "weight"),
               "decoder.conv_out.bias": (vae.decoder.conv_out, "bias")}
    with safe_open(path, framework="pt") as f:
        keys = set(f.keys())
        for key, (mod, attr) in targets.items():
            if key not in keys:
                continue
            t = f.get_tensor(key)
            cur = getattr(mod, attr)
            if tuple(t.shape) != tuple(cur.shape):
                continue
            if torch.isnan(t).any() or torch.isinf(t).any():
                raise RuntimeError(f"{key} contains NaN/Inf after loading RGBA checkpoint.")
            with torch.no_grad():
                cur.copy_(t.to(device=cur.device, dtype=cur.dtype))
            restored = True
    return restored


class RgbaVAE(nn.Module):
    """src/models/rgba_vae.py:194-342.  ``forward(x)`` takes (B,3|4,H,W) in [0,1] and returns
    ``(recon in [0,1], posterior)``; ``noise=`` makes the posterior sample reproducible."""

    def __init__(self, vae: RgbaAutoencoder, loss_reduce_mean: bool = False, use_naive_mse: bool = False,
                 custom_eb: Optional[Sequence[float]] = None, custom_eb2: Optional[Sequence[float]] = None, beta: float = 0.25,
                 **legacy_weights):
        super().__init__()
        from .losses import AlphaVaeLoss

        self.vae = vae
        self.beta = beta
        self.legacy_weights = legacy_weights  # alpha_loss_weight etc. of the reference's pre-AlphaVAE loss
        self._graphs = {}
        self.loss_module = AlphaVaeLoss(reduce_mean=loss_reduce_mean, use_naive_mse=use_naive_mse, custom_eb=custom_eb,
                                        custom_eb2=custom_eb2)

    @classmethod
    def from_pretrained_rgb(cls, model_name_or_path: str, subfolder: Optional[str] = "vae",
                            torch_dtype: Optional[torch.dtype] = torch.float32, alpha_bias_init: float = 0.0,
                            device: Optional[torch.device] = None, **kwargs) -> "RgbaVAE":
        if subfolder and not os.path.isfile(os.path.join(model_name_or_path, subfolder, CONFIG_NAME)):
            subfolder = None
        vae = RgbaAutoencoder.from_pretrained(model_name_or_path, subfolder=subfolder, torch_dtype=torch_dtype,
                                              ignore_mismatched_sizes=True, low_cpu_mem_usage=False)
        adapt_vae_to_rgba(vae, alpha_bias_init=alpha_bias_init)
        _maybe_restore_rgba_convs(vae, model_name_or_path, subfolder)
        if device is not None:
            vae = vae.to(device)
        return cls(vae=vae, **kwargs)

    def forward(self, x: torch.Tensor, noise: Optional[torch.Tensor] = None, generator=None):
        x_rgba = _ensure_alpha(x)
        # _to_vae_range is the conv_in loader's affine; _from_vae_range + clamp(0,1) the conv_out epilogue's
        moments = self.vae._run_sliced(lambda t: self.vae._encode_moments(t, in_scale=2.0, in_shift=-1.0), x_rgba)
        from .posterior import DiagonalGaussianDistribution

        posterior = DiagonalGaussianDistribution(moments)
        z = posterior.sample(generator=generator, noise=noise)
        recon = self.vae._run_sliced(
            lambda t: self.vae._decode_image(t, out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0)), z)
        return recon, posterior

    @torch.no_grad()
    def reconstruct(self, x: torch.Tensor) -> torch.Tensor:
        return self.forward(x)[0]

    # ---- CUDA-graph replay of the inference step -------------------------------------------
    @torch.no_grad()
    def forward_graphed(self, x: torch.Tensor, noise: torch.Tensor, backgrounds=((1.0, 1.0, 1.0),)):
        """``forward`` + validation metrics as ONE captured CUDA graph per (shape, dtype): ~130 kernel launches
        replay without host work in between.  Returns ``(recon, moments, metrics)`` -- views of the graph's static
        output buffers, valid until the next call with the same shape.  Weights must not change between calls
        (the packed-weight cache is baked into the graph); ``reset_graphs()`` drops the captures."""
        x = _ensure_alpha(x)
        key = (tuple(x.shape), x.dtype, tuple(noise.shape), noise.dtype, tuple(tuple(b) for b in backgrounds))
        g = self._graphs.get(key)
        if g is None:
            sx, sn = x.clone(), noise.clone()
            stream = torch.cuda.Stream()
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                for _ in range(2):  # warm-up: weight packing, allocator pools, TMA attribute set-up
                    recon, post = self.forward(sx, noise=sn)
                    ops.composite_psnr(recon, sx, backgrounds)
            torch.cuda.current_stream().wait_stream(stream)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                recon, post = self.forward(sx, noise=sn)
                metrics = ops.composite_psnr(recon, sx, backgrounds)
            g = self._graphs[key] = (graph, sx, sn, recon, post.parameters, metrics)
        graph, sx, sn, recon, moments, metrics = g
        sx.copy_(x, non_blocking=True)
        sn.copy_(noise, non_blocking=True)
        graph.replay()
        return recon, moments, metrics

    def reset_graphs(self) -> None:
        self._graphs.clear()

    def loss(self, recon: torch.Tensor, target: torch.Tensor, posterior) -> torch.Tensor:
        """AlphaVAE reconstruction term + beta * KL (rgba_vae.py:283-316 with the default weights of
        configs/flux_vae.yaml; inputs in [0,1])."""
        rec = self.loss_module.reconstruction_loss(_to_vae_range(_ensure_alpha(recon)), _to_vae_range(_ensure_alpha(target)))
        return rec + self.beta * posterior.kl().mean()
