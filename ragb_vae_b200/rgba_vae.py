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


def _locate_weight_file(directory: str) -> Optional[str]:
    """rgba_vae.py:135-140 (plus diffusers' own name for a pickled checkpoint)."""
    for filename in (WEIGHTS_NAME, "pytorch_model.bin", "diffusion_pytorch_model.bin"):
        candidate = os.path.join(directory, filename)
        if os.path.isfile(candidate):
            return candidate
    return None


def _maybe_restore_rgba_convs(vae, model_name_or_path: str, subfolder: Optional[str]) -> bool:
    """rgba_vae.py:143-191 -- a checkpoint saved after widening still says ``in_channels: 3`` in its
    config while its tensors are 4-channel; re-read ``encoder.conv_in`` / ``decoder.conv_out`` (weight and bias)
    straight from the weight file (``.safetensors``, or a pickled ``.bin`` state dict loaded with
    ``weights_only=True``).  Silently does nothing when the file or the tensors are missing or are not 4-channel,
    like the reference; returns True if the tensors were restored; raises RuntimeError on NaN/Inf."""
    root = str(model_name_or_path)
    if not os.path.exists(root):
        return False
    if subfolder:
        root = os.path.join(root, subfolder)
    if not os.path.isdir(root):
        return False
    path = _locate_weight_file(root)
    if path is None:
        return False
    wanted = ("encoder.conv_in.weight", "encoder.conv_in.bias", "decoder.conv_out.weight", "decoder.conv_out.bias")
    try:
        if path.endswith(".safetensors"):
            from safetensors import safe_open

            with safe_open(path, framework="pt") as f:
                keys = set(f.keys())
                sd = {k: f.get_tensor(k) for k in wanted if k in keys}
        else:
            full = torch.load(path, map_location="cpu", weights_only=True)
            sd = {k: full[k] for k in wanted if k in full}
    except Exception:
        return False
    w_in, w_out = sd.get("encoder.conv_in.weight"), sd.get("decoder.conv_out.weight")
    if w_in is None or w_out is None or w_in.shape[1] != 4 or w_out.shape[0] != 4:
        return False
    conv_in, conv_out = vae.encoder.conv_in, vae.decoder.conv_out
    with torch.no_grad():
        for mod, attr, key in ((conv_in, "weight", "encoder.conv_in.weight"), (conv_in, "bias", "encoder.conv_in.bias"),
                               (conv_out, "weight", "decoder.conv_out.weight"), (conv_out, "bias", "decoder.conv_out.bias")):
            cur, t = getattr(mod, attr, None), sd.get(key)
            if cur is not None and t is not None:
                cur.copy_(t.to(device=cur.device, dtype=cur.dtype))
    for name, tensor in (("encoder.conv_in.weight", conv_in.weight), ("decoder.conv_out.weight", conv_out.weight)):
        if torch.isnan(tensor).any() or torch.isinf(tensor).any():
            raise RuntimeError(f"{name} contains NaN/Inf after loading RGBA checkpoint.")
    return True


class RgbaVAE(nn.Module):
    """src/models/rgba_vae.py:194-342.  ``forward(x)`` takes (B,3|4,H,W) in [0,1] and returns
    ``(recon in [0,1], posterior)``; ``noise=`` makes the posterior sample reproducible."""

    def __init__(self, vae: RgbaAutoencoder, beta: float = 0.25, alpha_loss_weight: float = 1.0, alpha_l1_weight: float = 0.0,
                 rgb_loss_weight: float = 1.0, white_bg_weight: float = 0.0, black_bg_weight: float = 0.0,
                 loss_reduce_mean: bool = False, use_naive_mse: bool = False, custom_eb: Optional[Sequence[float]] = None,
                 custom_eb2: Optional[Sequence[float]] = None) -> None:
        """Same arguments, order and defaults as rgba_vae.py:195-208."""
        super().__init__()
        from .losses import AlphaVaeLoss

        self.vae = vae
        self.beta = beta
        self.alpha_loss_weight = alpha_loss_weight
        self.alpha_l1_weight = alpha_l1_weight
        self.rgb_loss_weight = rgb_loss_weight
        self.white_bg_weight = white_bg_weight
        self.black_bg_weight = black_bg_weight
        self.loss_reduce_mean = loss_reduce_mean
        self.use_naive_mse = use_naive_mse
        self._graphs = {}
        self._graphs_generation = 0
        # validates custom_eb / custom_eb2 like rgba_vae.py:224-225 and carries the Eb constants
        self.loss_module = AlphaVaeLoss(reduce_mean=loss_reduce_mean, use_naive_mse=use_naive_mse, custom_eb=custom_eb,
                                        custom_eb2=custom_eb2)
        self.register_buffer("alphavae_eb", self.loss_module.eb.clone(), persistent=False)
        self.register_buffer("alphavae_eb2", self.loss_module.eb2.clone(), persistent=False)

    @classmethod
    def from_pretrained_rgb(cls, model_name_or_path: str, subfolder: Optional[str] = "vae",
                            torch_dtype: Optional[torch.dtype] = torch.float32, alpha_bias_init: float = 0.0,
                            beta: float = 0.25, alpha_loss_weight: float = 1.0, alpha_l1_weight: float = 0.0,
                            rgb_loss_weight: float = 1.0, white_bg_weight: float = 0.0, black_bg_weight: float = 0.0,
                            device: Optional[torch.device] = None, loss_reduce_mean: bool = False, use_naive_mse: bool = False,
                            custom_eb: Optional[Sequence[float]] = None, custom_eb2: Optional[Sequence[float]] = None) -> "RgbaVAE":
        """rgba_vae.py:231-272.  One convenience beyond the reference: a checkpoint directory without the ``vae/``
        sub-folder (what ``save_pretrained`` writes) is accepted with the default ``subfolder``."""
        if subfolder and not os.path.isfile(os.path.join(model_name_or_path, subfolder, CONFIG_NAME)) \
                and os.path.isfile(os.path.join(model_name_or_path, CONFIG_NAME)):
            subfolder = None
        vae = RgbaAutoencoder.from_pretrained(model_name_or_path, subfolder=subfolder, torch_dtype=torch_dtype,
                                              ignore_mismatched_sizes=True, low_cpu_mem_usage=False)
        adapt_vae_to_rgba(vae, alpha_bias_init=alpha_bias_init)
        _maybe_restore_rgba_convs(vae, model_name_or_path, subfolder)
        if device is not None:
            vae = vae.to(device)
        return cls(vae=vae, beta=beta, alpha_loss_weight=alpha_loss_weight, alpha_l1_weight=alpha_l1_weight,
                   rgb_loss_weight=rgb_loss_weight, white_bg_weight=white_bg_weight, black_bg_weight=black_bg_weight,
                   loss_reduce_mean=loss_reduce_mean, use_naive_mse=use_naive_mse, custom_eb=custom_eb, custom_eb2=custom_eb2)

    def forward(self, x: torch.Tensor, noise: Optional[torch.Tensor] = None, generator=None):
        """rgba_vae.py:274-281.  Goes through the same tiled / sliced encode and decode as ``vae.encode`` /
        ``vae.decode`` (the reference stage enables both, rgba_vae_stage.py:296-304), so ``model(x)`` and
        ``model.vae.encode(x)`` always agree; ``_to_vae_range`` is the conv_in loader's affine and
        ``_from_vae_range`` + ``clamp(0, 1)`` the conv_out epilogue's (after the seam blend when tiled)."""
        from .posterior import DiagonalGaussianDistribution

        x_rgba = _ensure_alpha(x)
        vae = self.vae
        moments = vae._run_sliced(lambda t: vae._encode_maybe_tiled(t, in_scale=2.0, in_shift=-1.0), x_rgba)
        posterior = DiagonalGaussianDistribution(moments)
        z = posterior.sample(generator=generator, noise=noise)
        recon = vae._run_sliced(lambda t: vae._decode_maybe_tiled(t, out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0)), z)
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
        if self._graphs and self._graphs_generation != self.vae.weights_generation:
            self._graphs.clear()  # captured with packed copies of weights that have been updated since
        self._graphs_generation = self.vae.weights_generation
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
        """rgba_vae.py:283-316: ``rgb_loss_weight`` x (AlphaVAE map, or rgb MSE in [0,1] with ``use_naive_mse``) +
        ``white_bg_weight`` / ``black_bg_weight`` x MSE of the composites + ``alpha_loss_weight`` x alpha MSE +
        ``alpha_l1_weight`` x alpha L1 + ``beta`` x mean KL.  recon / target (B,3|4,H,W) in [0,1].  All pixel terms
        come from ONE pass over the pair (``rv_rgba_loss_terms``; the reference runs ~40 elementwise kernels).
        Difference: with ``use_naive_mse`` and the per-sample-sum reduction the reference raises (``.view`` of a
        channel slice, rgba_vae.py:321); here that combination returns the value ``.reshape`` would give."""
        sums = ops.rgba_loss_terms(_ensure_alpha(recon), _ensure_alpha(target), self.loss_module._eb, self.loss_module._eb2)
        b, hw = sums.shape[0], float(recon.shape[2] * recon.shape[3])
        mean3 = lambda col: sums[:, col].sum() / (b * 3.0 * hw)   # F.mse_loss over (B,3,H,W)
        mean1 = lambda col: sums[:, col].sum() / (b * hw)         # over the alpha plane
        total = self.beta * posterior.kl().mean()
        if self.rgb_loss_weight > 0.0:
            col = 1 if self.use_naive_mse else 0
            base = mean3(col) if self.loss_reduce_mean else sums[:, col].mean()   # _reduce_loss, rgba_vae.py:318-322
            total = total + self.rgb_loss_weight * base
        if self.white_bg_weight > 0.0:
            total = total + self.white_bg_weight * mean3(2)
        if self.black_bg_weight > 0.0:
            total = total + self.black_bg_weight * mean3(3)
        if self.alpha_loss_weight > 0.0:
            total = total + self.alpha_loss_weight * mean1(4)
        if self.alpha_l1_weight > 0.0:
            total = total + self.alpha_l1_weight * mean1(5)
        return total
