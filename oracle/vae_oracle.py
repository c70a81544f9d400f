"""CPU oracle for the RGBA-VAE hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain PyTorch fp32 restatement of the arithmetic the reference
executes below its drop-in boundary.  It is used only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs, as the checker.  Nothing under ``ragb_vae_b200/`` imports it.

PARITY, what is pinned and what is not.  The reference (jaejung-dev/ragb-vae) ships no tests, golden vectors or
fixtures.  Its OWN pure-torch functions restated here (losses, composites, PSNR, triplet,
blend, batch assembly, adapt_vae_to_rgba, RgbaVAE.forward / .loss, the validation loop) ARE pinned to the reference
itself: scripts/make_reference_fixtures.py executes the unmodified reference files through the sys.modules shim of
tests/refshim.py and commits their outputs under tests/golden/ref_*; tests/test_reference_pin.py holds this file
to them (and re-derives them live where /root/reference exists).  PARITY UNPINNED for the diffusers-internal half:
the conv / norm / attention stacks and the posterior live in the un-vendored, un-pinned third-party package
``diffusers`` (requirements.txt:2), which is not installable here.  That half follows the published diffusers
algorithms:

* ``arch="flux"``  -- ``diffusers.AutoencoderKL`` with the FLUX.1 vae config
  (``models/autoencoders/autoencoder_kl.py``, ``vae.py::Encoder/Decoder``,
  ``resnet.py::ResnetBlock2D``, ``attention_processor.py::AttnProcessor2_0``).
* ``arch="qwen"``  -- ``diffusers.AutoencoderKLQwenImage``
  (``models/autoencoders/autoencoder_kl_qwenimage.py``) evaluated on a single
  frame (T=1), accepting 4-D tensors like the reference call sites do
  (src/models/rgba_vae.py:275-279).

Anchors that ARE checked (tests/test_oracle.py): exact public parameter counts
(Flux 83 819 683 / Qwen 126 892 531 for 3 channels; 83 821 988 / 126 897 716
after RGBA widening), bit-level agreement of the flux oracle with the
independent BFL/torchtitan auto-encoder that ships in this image (vectors
committed under tests/golden/ by scripts/make_golden.py), and the identity
causal-conv3d(T=1) == conv2d(w[:, :, 2]).

Reference call sites restated here:
  src/models/rgba_vae.py:25-37   _ensure_alpha / _to_vae_range / _from_vae_range
  src/models/rgba_vae.py:40-92   _normalize_background / composite_over_*
  src/models/rgba_vae.py:95-123  adapt_vae_to_rgba
  src/models/rgba_vae.py:274-281 RgbaVAE.forward
  src/models/losses.py:67-83,109-123  reconstruction_loss / kl_loss / _reduce
  src/training/rgba_vae_stage.py:606-625,690-700,712-715  triplet / split / psnr
  src/training/rgba_vae_stage.py:433-518  the train step up to backward (training_step; torch autograd gives the gradients)
  src/training/rgba_vae_stage.py:85-130,575-603  RandomBackgroundBlend._blend_tensor / build_training_batch
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# DiagonalGaussianDistribution (diffusers models/autoencoders/vae.py)
# --------------------------------------------------------------------------- #
class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator=None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        # diffusers draws randn_tensor(mean.shape, generator, device, dtype); a supplied
        # noise tensor is our extension so that parity is definable.
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype, device=self.parameters.device)
        return self.mean + self.std * noise.to(self.mean.dtype)

    def kl(self, other: "Optional[DiagonalGaussianDistribution]" = None) -> torch.Tensor:
        if self.deterministic:
            return torch.tensor([0.0], device=self.parameters.device)
        if other is None:
            return 0.5 * torch.sum(self.mean.pow(2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(
            (self.mean - other.mean).pow(2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
            dim=[1, 2, 3],
        )

    def mode(self) -> torch.Tensor:
        return self.mean


# --------------------------------------------------------------------------- #
# arch = "flux": diffusers.AutoencoderKL
# --------------------------------------------------------------------------- #
class FluxResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class FluxAttention(nn.Module):
    """diffusers Attention with AttnProcessor2_0, one head, residual connection."""

    def __init__(self, c: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps, affine=True)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        residual = x
        t = x.view(b, c, h * w).transpose(1, 2)
        t = self.group_norm(t.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, h, w) + residual


class _ConvHolder(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class FluxDownBlock(nn.Module):
    def __init__(self, cin, cout, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([FluxResnetBlock2D(cin, cout), FluxResnetBlock2D(cout, cout)])
        self.downsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, stride=2, padding=0))]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0].conv(F.pad(x, (0, 1, 0, 1)))
        return x


class FluxUpBlock(nn.Module):
    def __init__(self, cin, cout, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([FluxResnetBlock2D(cin if i == 0 else cout, cout) for i in range(3)])
        self.upsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, padding=1))]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0].conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class FluxMidBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attentions = nn.ModuleList([FluxAttention(c)])
        self.resnets = nn.ModuleList([FluxResnetBlock2D(c, c), FluxResnetBlock2D(c, c)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class FluxEncoder(nn.Module):
    def __init__(self, in_channels, latent_channels, boc):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        blocks, c = [], boc[0]
        for i, co in enumerate(boc):
            blocks.append(FluxDownBlock(c, co, add_down=i != len(boc) - 1))
            c = co
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = FluxMidBlock(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class FluxDecoder(nn.Module):
    def __init__(self, out_channels, latent_channels, boc):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(latent_channels, rev[0], 3, padding=1)
        self.mid_block = FluxMidBlock(rev[0])
        blocks, c = [], rev[0]
        for i, co in enumerate(rev):
            blocks.append(FluxUpBlock(c, co, add_up=i != len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


# --------------------------------------------------------------------------- #
# arch = "qwen": diffusers.AutoencoderKLQwenImage at T = 1
# --------------------------------------------------------------------------- #
class QwenCausalConv3d(nn.Conv3d):
    """nn.Conv3d with all temporal padding in front (F.pad(x, (pw,pw,ph,ph,2*pt,0)))."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0):
        super().__init__(cin, cout, kernel_size, stride=stride, padding=0)
        p = (padding,) * 3 if isinstance(padding, int) else tuple(padding)
        self._pad = (p[2], p[2], p[1], p[1], 2 * p[0], 0)

    def forward(self, x):  # x: (B, C, T, H, W) -- the literal form
        return super().forward(F.pad(x, self._pad))

    def forward_frame(self, x):  # x: (B, C, H, W), T = 1: only the last temporal tap sees data
        kt = self.weight.shape[2]
        return F.conv2d(x, self.weight[:, :, kt - 1], self.bias, padding=(self._pad[2], self._pad[0]))


class QwenRMSNorm(nn.Module):
    def __init__(self, dim, images=True):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones((dim, 1, 1) if images else (dim, 1, 1, 1)))

    def forward(self, x):  # x: (B, C, H, W)
        return F.normalize(x, dim=1) * self.scale * self.gamma.reshape(1, -1, 1, 1)


class QwenResidualBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = QwenRMSNorm(cin, images=False)
        self.conv1 = QwenCausalConv3d(cin, cout, 3, padding=1)
        self.norm2 = QwenRMSNorm(cout, images=False)
        self.conv2 = QwenCausalConv3d(cout, cout, 3, padding=1)
        self.conv_shortcut = QwenCausalConv3d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = x if self.conv_shortcut is None else self.conv_shortcut.forward_frame(x)
        x = self.conv1.forward_frame(F.silu(self.norm1(x)))
        x = self.conv2.forward_frame(F.silu(self.norm2(x)))
        return x + h


class QwenAttentionBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = QwenRMSNorm(dim, images=True)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)

    def forward(self, x):
        b, c, h, w = x.shape
        qkv = self.to_qkv(self.norm(x)).reshape(b, 1, c * 3, -1).permute(0, 1, 3, 2).contiguous()
        q, k, v = qkv.chunk(3, dim=-1)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.squeeze(1).permute(0, 2, 1).reshape(b, c, h, w)
        return self.proj(o) + x


class QwenResample(nn.Module):
    def __init__(self, dim, mode):
        super().__init__()
        self.mode = mode
        if mode in ("upsample2d", "upsample3d"):
            self.resample = nn.Sequential(nn.Identity(), nn.Conv2d(dim, dim // 2, 3, padding=1))
            if mode == "upsample3d":
                self.time_conv = QwenCausalConv3d(dim, dim * 2, (3, 1, 1), padding=(1, 0, 0))  # video only
        else:
            self.resample = nn.Sequential(nn.ZeroPad2d((0, 1, 0, 1)), nn.Conv2d(dim, dim, 3, stride=2))
            if mode == "downsample3d":
                self.time_conv = QwenCausalConv3d(dim, dim, (3, 1, 1), stride=(2, 1, 1), padding=(0, 0, 0))  # video only

    def forward(self, x):
        if self.mode.startswith("upsample"):
            x = F.interpolate(x.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(x)
            return self.resample[1](x)
        return self.resample(x)


class QwenMidBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.resnets = nn.ModuleList([QwenResidualBlock(dim, dim), QwenResidualBlock(dim, dim)])
        self.attentions = nn.ModuleList([QwenAttentionBlock(dim)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class QwenEncoder3d(nn.Module):
    def __init__(self, in_channels=4, dim=96, z_dim=32, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_downsample=(False, True, True)):
        super().__init__()
        dims = [dim * u for u in (1,) + tuple(dim_mult)]
        self.conv_in = QwenCausalConv3d(in_channels, dims[0], 3, padding=1)
        blocks = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                blocks.append(QwenResidualBlock(cin, cout))
                cin = cout
            if i != len(dim_mult) - 1:
                blocks.append(QwenResample(cout, "downsample3d" if temperal_downsample[i] else "downsample2d"))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = QwenMidBlock(dims[-1])
        self.norm_out = QwenRMSNorm(dims[-1], images=False)
        self.conv_out = QwenCausalConv3d(dims[-1], z_dim, 3, padding=1)

    def forward(self, x):
        x = self.conv_in.forward_frame(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out.forward_frame(F.silu(self.norm_out(x)))


class QwenUpBlock(nn.Module):
    def __init__(self, cin, cout, num_res_blocks, upsample_mode):
        super().__init__()
        self.resnets = nn.ModuleList(
            [QwenResidualBlock(cin if i == 0 else cout, cout) for i in range(num_res_blocks + 1)])
        self.upsamplers = nn.ModuleList([QwenResample(cout, upsample_mode)]) if upsample_mode else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class QwenDecoder3d(nn.Module):
    def __init__(self, out_channels=4, dim=96, z_dim=16, dim_mult=(1, 2, 4, 4), num_res_blocks=2,
                 temperal_upsample=(True, True, False)):
        super().__init__()
        dims = [dim * u for u in (dim_mult[-1],) + tuple(dim_mult[::-1])]
        self.conv_in = QwenCausalConv3d(z_dim, dims[0], 3, padding=1)
        self.mid_block = QwenMidBlock(dims[0])
        blocks = []
        for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
            if i > 0:
                cin = cin // 2
            mode = None
            if i != len(dim_mult) - 1:
                mode = "upsample3d" if temperal_upsample[i] else "upsample2d"
            blocks.append(QwenUpBlock(cin, cout, num_res_blocks, mode))
        self.up_blocks = nn.ModuleList(blocks)
        self.norm_out = QwenRMSNorm(dims[-1], images=False)
        self.conv_out = QwenCausalConv3d(dims[-1], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in.forward_frame(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out.forward_frame(F.silu(self.norm_out(x)))


# --------------------------------------------------------------------------- #
# Model-level surface (encode / decode), both arches
# --------------------------------------------------------------------------- #
class OracleVAE(nn.Module):
    """``vae.encode(x).latent_dist`` / ``vae.decode(z).sample`` in fp32 on the CPU."""

    def __init__(self, arch: str = "qwen", in_channels: int = 4, out_channels: int = 4):
        super().__init__()
        self.arch = arch
        if arch == "flux":
            boc = (128, 256, 512, 512)
            self.encoder = FluxEncoder(in_channels, 16, boc)
            self.decoder = FluxDecoder(out_channels, 16, boc)
            self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, latent_channels=16,
                                          block_out_channels=list(boc), scaling_factor=0.3611, shift_factor=0.1159,
                                          sample_size=1024)
        elif arch == "qwen":
            self.encoder = QwenEncoder3d(in_channels, 96, 32)
            self.quant_conv = QwenCausalConv3d(32, 32, 1)
            self.post_quant_conv = QwenCausalConv3d(16, 16, 1)
            self.decoder = QwenDecoder3d(out_channels, 96, 16)
            self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, z_dim=16, base_dim=96,
                                          latent_channels=16, block_out_channels=[96, 192, 384, 384], sample_size=256)
        else:
            raise ValueError(f"unknown arch {arch!r}")

    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        h = self.encoder(x)
        if self.arch == "qwen":
            h = self.quant_conv.forward_frame(h)
        return h

    def encode(self, x: torch.Tensor):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.encode_moments(x)))

    def decode(self, z: torch.Tensor):
        if self.arch == "qwen":
            y = self.decoder(self.post_quant_conv.forward_frame(z))
            y = torch.clamp(y, min=-1.0, max=1.0)  # AutoencoderKLQwenImage._decode
        else:
            y = self.decoder(z)
        return SimpleNamespace(sample=y)


def build_oracle(arch: str, seed: int = 0, rgba_random: bool = True) -> OracleVAE:
    """Random-init (PyTorch default init) under ``torch.manual_seed(seed)``.

    The model is constructed directly with 4 in/out channels so that the alpha
    column/row carry random (non-zero) weights and the alpha path is exercised
    (SURVEY 8d); ``rgba_random=False`` instead builds the 3-channel model and
    widens it exactly like ``adapt_vae_to_rgba`` (zero alpha column / row).
    """
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        if rgba_random:
            m = OracleVAE(arch, 4, 4)
        else:
            m = OracleVAE(arch, 3, 3)
            adapt_vae_to_rgba(m)
    finally:
        torch.random.set_rng_state(gen_state)
    return m.eval().requires_grad_(False)


# --------------------------------------------------------------------------- #
# In-repo reference functions restated (pure torch)
# --------------------------------------------------------------------------- #
def ensure_alpha(x):  # src/models/rgba_vae.py:25-29
    if x.shape[1] == 4:
        return x
    return torch.cat([x, torch.ones((x.shape[0], 1, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device)], dim=1)


def to_vae_range(x):  # :32-33
    return x * 2.0 - 1.0


def from_vae_range(x):  # :36-37
    return (x + 1.0) * 0.5


def normalize_background(background, reference):  # :40-72
    dtype = reference.dtype
    batch, _, height, width = reference.shape
    if isinstance(background, torch.Tensor):
        bg = background.to(device=reference.device, dtype=dtype)
        if bg.dim() == 3:
            bg = bg.unsqueeze(0)
        if bg.dim() != 4:
            raise ValueError(f"Background tensor must have 3 or 4 dimensions, got {bg.dim()}")
        if bg.shape[0] == 1 and batch > 1:
            bg = bg.expand(batch, -1, -1, -1)
        if bg.shape[1] == 1:
            bg = bg.repeat(1, 3, 1, 1)
        if bg.shape[2] != height or bg.shape[3] != width:
            raise ValueError("Background tensor spatial size must match the RGBA tensor.")
        return bg
    if isinstance(background, Sequence):
        if len(background) != 3:
            raise ValueError("Background color sequence must contain exactly three values.")
        return torch.tensor(background, dtype=dtype, device=reference.device).view(1, 3, 1, 1).expand(batch, -1, height, width)
    return torch.full((batch, 3, height, width), float(background), dtype=dtype, device=reference.device)


def composite_over_background(rgba, background):  # :75-84
    rgba = ensure_alpha(rgba)
    rgb, alpha = rgba[:, :3], rgba[:, 3:4]
    return rgb * alpha + normalize_background(background, rgb) * (1.0 - alpha)


def adapt_vae_to_rgba(vae, alpha_bias_init: float = 0.0) -> None:  # :95-123 (rank-agnostic)
    conv_in = vae.encoder.conv_in
    if conv_in.in_channels != 4:
        w = conv_in.weight.data
        nw = torch.zeros(w.size(0), 4, *w.shape[2:], dtype=w.dtype, device=w.device)
        nw[:, :3] = w
        conv_in.in_channels = 4
        conv_in.weight = nn.Parameter(nw)
    conv_out = vae.decoder.conv_out
    if conv_out.out_channels != 4:
        w = conv_out.weight.data
        nw = torch.zeros(4, w.size(1), *w.shape[2:], dtype=w.dtype, device=w.device)
        nw[:3] = w
        conv_out.out_channels = 4
        conv_out.weight = nn.Parameter(nw)
        nb = torch.zeros(4, dtype=w.dtype, device=w.device)
        nb[:3] = conv_out.bias.data
        nb[3] = alpha_bias_init
        conv_out.bias = nn.Parameter(nb)
    vae.config.in_channels = 4
    vae.config.out_channels = 4


def rgba_vae_forward(vae: OracleVAE, x: torch.Tensor, noise: torch.Tensor):
    """RgbaVAE.forward (src/models/rgba_vae.py:274-281) with a supplied noise tensor."""
    posterior = vae.encode(to_vae_range(ensure_alpha(x))).latent_dist
    z = posterior.sample(noise=noise)
    recon = torch.clamp(from_vae_range(vae.decode(z).sample), 0.0, 1.0)
    return recon, posterior, z


EB = (-0.0357, -0.0811, -0.1797)
EB2 = (0.3163, 0.3060, 0.3634)


def reduce_loss(value, reduce_mean: bool):  # src/models/losses.py:117-123
    if value.ndim == 0:
        return value
    if reduce_mean:
        return value.mean()
    return value.view(value.shape[0], -1).sum(dim=1).mean()


def reconstruction_loss(pred, target, reduce_mean=False, use_naive_mse=False, eb=EB, eb2=EB2):
    """AlphaVaeLoss.reconstruction_loss (src/models/losses.py:67-83); inputs in [-1, 1] RGBA."""
    if use_naive_mse:
        return reduce_loss((pred - target).pow(2), reduce_mean)
    eb_t = torch.tensor(eb, dtype=torch.float32, device=pred.device).view(1, 3, 1, 1)
    eb2_t = torch.tensor(eb2, dtype=torch.float32, device=pred.device).view(1, 3, 1, 1)
    ta = (target[:, 3:] + 1.0) * 0.5
    pa = (pred[:, 3:] + 1.0) * 0.5
    d = target[:, :3] * ta - pred[:, :3] * pa
    da = ta - pa
    return reduce_loss(d.pow(2) - 2.0 * eb_t * d * da + eb2_t * da.pow(2), reduce_mean)


def kl_loss(posterior, reference=None, reduce_mean=False):  # losses.py:109-115
    return reduce_loss(posterior.kl(reference), reduce_mean)


def training_step(vae: OracleVAE, inputs: torch.Tensor, noise: torch.Tensor, kl_scale: Optional[float] = 1e-6,
                  reduce_mean: bool = False, use_naive_mse: bool = False, ref_vae: Optional[OracleVAE] = None,
                  ref_kl_scale: Optional[float] = None):
    """One ``rgba_vae`` step up to ``accelerator.backward`` (src/training/rgba_vae_stage.py:433-518) with
    ``lpips_scale = 0`` (the reference-KL term is optional: ``ref_vae`` + ``ref_kl_scale``): returns (metrics, {parameter name: gradient}).  ``inputs`` in [0,1];
    ``noise`` is the posterior sample's eps for the first third of the triplet batch (reproducible ``sample()``)."""
    vae.requires_grad_(True)
    vae.zero_grad(set_to_none=True)
    target = torch.clamp(inputs, 0.0, 1.0)
    target_vae = target * 2.0 - 1.0
    composed = build_detail_augmented_triplet(target_vae)
    posterior_all = vae.encode(composed).latent_dist
    posterior, posterior_black, posterior_white = split_triplet_distribution(posterior_all)
    z = posterior.sample(noise=noise)
    pred = vae.decode(z).sample
    recon = reconstruction_loss(pred, target_vae, reduce_mean, use_naive_mse)
    metrics = {"train/recon": recon.detach()}
    total = recon
    if kl_scale is not None and kl_scale > 0.0:
        kl = kl_loss(posterior, None, reduce_mean)
        metrics["train/kl"] = kl.detach()
        total = total + kl_scale * kl
    if ref_vae is not None and ref_kl_scale and ref_kl_scale > 0.0:  # rgba_vae_stage.py:489-508
        with torch.no_grad():
            ref_all = ref_vae.encode(composed).latent_dist
        _, ref_black, ref_white = split_triplet_distribution(ref_all)
        ref_kl = 0.5 * (kl_loss(posterior_black, ref_black, reduce_mean) + kl_loss(posterior_white, ref_white, reduce_mean))
        metrics["train/ref_kl"] = ref_kl.detach()
        total = total + ref_kl_scale * ref_kl
    metrics["train/loss"] = total.detach()
    total.backward()
    grads = {n: p.grad.detach().clone() for n, p in vae.named_parameters() if p.grad is not None}
    vae.requires_grad_(False)
    return metrics, grads


def compute_psnr(pred, target):  # src/training/rgba_vae_stage.py:712-715
    mse = torch.clamp(torch.mean((pred - target) ** 2, dim=(1, 2, 3)), min=1e-8)
    return -10.0 * torch.log10(mse)


def alpha_mae(recon, inputs):  # rgba_vae_stage.py:749-753
    return torch.mean(torch.abs(recon[:, 3:] - inputs[:, 3:]), dim=(1, 2, 3))


def build_detail_augmented_triplet(target):  # rgba_vae_stage.py:606-625
    if target.shape[1] < 4:
        raise ValueError("detail augmentation expects RGBA tensors.")
    fg = (1.0 + target[:, 3:4]) * 0.5
    bg = (1.0 - target[:, 3:4]) * 0.5
    black = (target * fg - bg).clone()
    white = (target * fg + bg).clone()
    black[:, 3:] = 1.0
    white[:, 3:] = 1.0
    return torch.cat([target, black, white], dim=0)


def background_blend(tensor, color):  # RandomBackgroundBlend._blend_tensor, rgba_vae_stage.py:118-129 (colour supplied)
    """tensor: (4,H,W) RGBA in [0,1]; color: (3,) -- composite over the opaque colour, alpha := 1."""
    tensor = tensor.clone()
    rgb, alpha = tensor[:3], tensor[3:4]
    bg = color.to(tensor.dtype).view(3, 1, 1).expand_as(rgb)
    blended = rgb * alpha + bg * (1.0 - alpha)
    return torch.cat([blended, torch.ones_like(alpha)], dim=0)


def build_training_batch(batch, background_mask=None):  # rgba_vae_stage.py:575-603 (mask supplied instead of torch.rand)
    tensors = []
    if "component" in batch and "composite" in batch:
        tensors.extend([batch["component"], batch["composite"]])
    elif "composite" in batch:
        tensors.append(batch["composite"])
    else:
        raise ValueError("Batch must contain 'composite' tensor for training.")
    inputs = torch.cat(tensors, dim=0)
    if background_mask is not None and "background" in batch:
        background = batch["background"]
        if background.dim() == 3:
            background = background.unsqueeze(0)
        if background.shape[1] != 4:
            raise ValueError("Background tensor is expected to have 4 channels (RGBA).")
        if background_mask.any():
            inputs = torch.cat([inputs, background[background_mask]], dim=0)
    return inputs


def split_triplet_distribution(posterior):  # rgba_vae_stage.py:690-700
    chunks = torch.chunk(posterior.parameters, 3, dim=0)
    if len(chunks) != 3:
        raise ValueError("Posterior batch dimension must be divisible by 3 for triplet splits.")
    return tuple(DiagonalGaussianDistribution(c) for c in chunks)


def validation_metrics(recon, inputs, backgrounds=(1.0, 0.0)):
    """The per-batch body of evaluate_rgba_vae (rgba_vae_stage.py:742-753), fp32."""
    out = {}
    for bg in backgrounds:
        out[bg] = compute_psnr(composite_over_background(recon, bg), composite_over_background(inputs, bg))
    out["alpha_mae"] = alpha_mae(recon, inputs)
    return out


def synthetic_rgba(batch: int, h: int, w: int, seed: int = 1, structured: bool = False) -> torch.Tensor:
    """Seeded synthetic RGBA in [0,1].  ``structured`` = smooth radial alpha with
    ~30 % exact 0 and ~30 % exact 1, the rest a ramp (real layers are mostly transparent)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 4, h, w, generator=g)
    if structured:
        yy = torch.linspace(-1, 1, h).view(h, 1)
        xx = torch.linspace(-1, 1, w).view(1, w)
        r = torch.sqrt(yy * yy + xx * xx) / math.sqrt(2.0)
        a = torch.clamp((0.75 - r) / 0.4 - 0.35, 0.0, 1.0)
        x[:, 3] = a
    return x


# --------------------------------------------------------------------------- #
# Flux latent plumbing and image I/O conversions
# --------------------------------------------------------------------------- #
def pack_latents(latents):
    """diffusers FluxPipeline._pack_latents, as called by src/models/flux_kontext_textalpha.py:334-341."""
    b, c, h, w = latents.shape
    x = latents.view(b, c, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(b, (h // 2) * (w // 2), c * 4)


def unpack_latents(tokens, height, width, vae_scale_factor=8):
    """diffusers FluxPipeline._unpack_latents (flux_kontext_textalpha.py:343-349)."""
    b, _, f = tokens.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    x = tokens.view(b, h // 2, w // 2, f // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(b, f // 4, h, w)


def load_rgba_u8(arr_u8):
    """inference_rgba_flux.py:15-20 after PIL: uint8 (H,W,4) -> float CHW in [0,1]."""
    return arr_u8.float().permute(2, 0, 1) / 255.0


def save_rgba_u8(tensor):
    """inference_rgba_flux.py:23-26 before PIL: CHW in [0,1] -> uint8 (H,W,4), truncating."""
    return (tensor.clamp(0, 1).float().permute(1, 2, 0) * 255).to(torch.uint8)


# --------------------------------------------------------------------------- #
# Tiled encode / decode (diffusers enable_tiling)
# --------------------------------------------------------------------------- #
def _blend_v(a, b, extent):
    extent = min(a.shape[2], b.shape[2], extent)
    for y in range(extent):
        b[:, :, y, :] = a[:, :, -extent + y, :] * (1 - y / extent) + b[:, :, y, :] * (y / extent)
    return b


def _blend_h(a, b, extent):
    extent = min(a.shape[3], b.shape[3], extent)
    for x in range(extent):
        b[:, :, :, x] = a[:, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, x] * (x / extent)
    return b


def _tiled(t, fn, tile, stride, blend, limit):
    """Row-major, in-place blending exactly like diffusers' tiled_encode / tiled_decode loops."""
    rows = []
    for i in range(0, t.shape[2], stride):
        rows.append([fn(t[:, :, i:i + tile, j:j + tile]).clone() for j in range(0, t.shape[3], stride)])
    result_rows = []
    for i, row in enumerate(rows):
        result_row = []
        for j, tile_ in enumerate(row):
            if i > 0:
                tile_ = _blend_v(rows[i - 1][j], tile_, blend)
            if j > 0:
                tile_ = _blend_h(row[j - 1], tile_, blend)
            result_row.append(tile_[:, :, :limit, :limit])
        result_rows.append(torch.cat(result_row, dim=3))
    return torch.cat(result_rows, dim=2)


def tiling_params(vae: "OracleVAE"):
    """(sample tile, sample stride, latent tile, latent stride, enc blend, enc limit, dec blend, dec limit)."""
    if vae.arch == "flux":  # AutoencoderKL: tile = sample_size, tile_overlap_factor = 0.25
        ts = int(vae.config.sample_size)
        tl = int(ts / (2 ** (len(vae.config.block_out_channels) - 1)))
        eb, db = int(tl * 0.25), int(ts * 0.25)
        return ts, int(ts * 0.75), tl, int(tl * 0.75), eb, tl - eb, db, ts - db
    ts, ss = 256, 192          # AutoencoderKLQwenImage defaults
    tl, sl = ts // 8, ss // 8
    return ts, ss, tl, sl, tl - sl, sl, ts - ss, ss


def tiled_encode_moments(vae: "OracleVAE", x):
    ts, ss, tl, sl, eb, el, _, _ = tiling_params(vae)
    if x.shape[-1] <= ts and x.shape[-2] <= ts:
        return vae.encode_moments(x)
    m = _tiled(x, vae.encode_moments, ts, ss, eb, el)
    return m[:, :, :x.shape[2] // 8, :x.shape[3] // 8]


def tiled_decode(vae: "OracleVAE", z):
    ts, ss, tl, sl, _, _, db, dl = tiling_params(vae)
    if z.shape[-1] <= tl and z.shape[-2] <= tl:
        return vae.decode(z).sample
    if vae.arch == "qwen":  # tiled_decode returns the blended tiles without the clamp of _decode
        fn = lambda t: vae.decoder(vae.post_quant_conv.forward_frame(t))
    else:
        fn = lambda t: vae.decoder(t)
    y = _tiled(z, fn, tl, sl, db, dl)
    return y[:, :, :z.shape[2] * 8, :z.shape[3] * 8]
