"""RgbaAutoencoder -- the drop-in for the diffusers VAE object the reference calls.

The reference never touches VAE internals beyond the duck-type listed in SURVEY.md 8(b):
``encode(x).latent_dist`` (src/models/rgba_vae.py:277, src/training/rgba_vae_stage.py:449,494,
src/models/flux_kontext_textalpha.py:331), ``decode(z).sample`` (rgba_vae.py:279,
rgba_vae_stage.py:452, flux_kontext_textalpha.py:497), ``config.*``, ``encoder.conv_in`` /
``decoder.conv_out`` (rgba_vae.py:97-120), ``enable_tiling/slicing/gradient_checkpointing``
(rgba_vae_stage.py:296-307), ``from_pretrained`` / ``save_pretrained`` / ``state_dict``.

Two architectures sit behind that surface:
  * ``arch="qwen"`` -- diffusers ``AutoencoderKLQwenImage`` evaluated on one frame: causal-conv3d
    (only the last temporal tap is live at T=1), per-pixel RMS norm + SiLU, single-head mid attention.
  * ``arch="flux"`` -- diffusers ``AutoencoderKL`` with the FLUX.1 VAE config: GroupNorm(32) + SiLU.
Module and parameter names equal the diffusers ones, so ``state_dict()`` round-trips checkpoints.

The modules below only HOLD parameters.  All arithmetic is librgbavae (hand-written sm_100a
kernels, through ops.py): bf16 models run the tcgen05 implicit-GEMM path with NHWC bf16
activations; fp32 models (parity config c1) run the CUDA-core path.  There is no eager fallback.
"""
from __future__ import annotations

import json
import math
import os
import warnings
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._lib import RV_BF16, RV_F32, RvError
from .posterior import DiagonalGaussianDistribution

WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
CONFIG_NAME = "config.json"


# ------------------------------------------------------------------------------------------
# parameter holders (names = diffusers names)
# ------------------------------------------------------------------------------------------
class Conv(nn.Module):
    """nn.Conv2d ([cout,cin,k,k]) or QwenImageCausalConv3d ([cout,cin,kt,kh,kw]) parameters."""

    def __init__(self, cin: int, cout: int, k: int, stride: int = 1, kernel3d: Optional[Tuple[int, int, int]] = None):
        super().__init__()
        self.in_channels, self.out_channels, self.k, self.stride = cin, cout, k, stride
        shape = (cout, cin, k, k) if kernel3d is None else (cout, cin, *kernel3d)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = cin * int(torch.tensor(shape[2:]).prod())
        bound = 1.0 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)

    def weight2d(self) -> torch.Tensor:
        w = self.weight
        if w.dim() == 5:  # causal conv3d on one frame: all temporal padding is in front
            w = w[:, :, -1]
        return w


class Linear(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.in_features, self.out_features = cin, cout
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        nn.init.uniform_(self.bias, -1.0 / math.sqrt(cin), 1.0 / math.sqrt(cin))


class GroupNorm(nn.Module):
    def __init__(self, c: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.num_groups, self.eps = groups, eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class RMSNorm(nn.Module):
    def __init__(self, c: int, images: bool = True):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones((c, 1, 1) if images else (c, 1, 1, 1)))


class _Holder(nn.Module):
    pass


def _flux_resnet(cin, cout):
    m = _Holder()
    m.norm1, m.conv1 = GroupNorm(cin), Conv(cin, cout, 3)
    m.norm2, m.conv2 = GroupNorm(cout), Conv(cout, cout, 3)
    if cin != cout:
        m.conv_shortcut = Conv(cin, cout, 1)
    return m


def _flux_attention(c):
    m = _Holder()
    m.group_norm = GroupNorm(c)
    m.to_q, m.to_k, m.to_v = Linear(c, c), Linear(c, c), Linear(c, c)
    m.to_out = nn.ModuleList([Linear(c, c), nn.Identity()])
    return m


def _flux_mid(c):
    m = _Holder()
    m.attentions = nn.ModuleList([_flux_attention(c)])
    m.resnets = nn.ModuleList([_flux_resnet(c, c), _flux_resnet(c, c)])
    return m


def _conv_holder(conv):
    m = _Holder()
    m.conv = conv
    return m


def _build_flux(in_ch, out_ch, latent, boc):
    enc = _Holder()
    enc.conv_in = Conv(in_ch, boc[0], 3)
    blocks, c = [], boc[0]
    for i, co in enumerate(boc):
        b = _Holder()
        b.resnets = nn.ModuleList([_flux_resnet(c, co), _flux_resnet(co, co)])
        if i != len(boc) - 1:
            b.downsamplers = nn.ModuleList([_conv_holder(Conv(co, co, 3, stride=2))])
        blocks.append(b)
        c = co
    enc.down_blocks = nn.ModuleList(blocks)
    enc.mid_block = _flux_mid(c)
    enc.conv_norm_out = GroupNorm(c)
    enc.conv_out = Conv(c, 2 * latent, 3)

    dec = _Holder()
    rev = list(reversed(boc))
    dec.conv_in = Conv(latent, rev[0], 3)
    dec.mid_block = _flux_mid(rev[0])
    blocks, c = [], rev[0]
    for i, co in enumerate(rev):
        b = _Holder()
        b.resnets = nn.ModuleList([_flux_resnet(c if j == 0 else co, co) for j in range(3)])
        if i != len(rev) - 1:
            b.upsamplers = nn.ModuleList([_conv_holder(Conv(co, co, 3))])
        blocks.append(b)
        c = co
    dec.up_blocks = nn.ModuleList(blocks)
    dec.conv_norm_out = GroupNorm(c)
    dec.conv_out = Conv(c, out_ch, 3)
    return enc, dec


def _qwen_resblock(cin, cout):
    m = _Holder()
    m.norm1, m.conv1 = RMSNorm(cin, images=False), Conv(cin, cout, 3, kernel3d=(3, 3, 3))
    m.norm2, m.conv2 = RMSNorm(cout, images=False), Conv(cout, cout, 3, kernel3d=(3, 3, 3))
    if cin != cout:
        m.conv_shortcut = Conv(cin, cout, 1, kernel3d=(1, 1, 1))
    m._kind = "res"
    return m


def _qwen_attention(c):
    m = _Holder()
    m.norm = RMSNorm(c, images=True)
    m.to_qkv = Conv(c, 3 * c, 1)
    m.proj = Conv(c, c, 1)
    return m


def _qwen_mid(c):
    m = _Holder()
    m.resnets = nn.ModuleList([_qwen_resblock(c, c), _qwen_resblock(c, c)])
    m.attentions = nn.ModuleList([_qwen_attention(c)])
    return m


def _qwen_resample(dim, mode):
    m = _Holder()
    m._kind, m.mode = "resample", mode
    if mode.startswith("upsample"):
        m.resample = nn.Sequential(nn.Identity(), Conv(dim, dim // 2, 3))
        if mode == "upsample3d":  # video-only temporal conv: never executed for one frame, kept for checkpoints
            m.time_conv = Conv(dim, dim * 2, 1, kernel3d=(3, 1, 1))
    else:
        m.resample = nn.Sequential(nn.Identity(), Conv(dim, dim, 3, stride=2))
        if mode == "downsample3d":
            m.time_conv = Conv(dim, dim, 1, kernel3d=(3, 1, 1))
    return m


def _build_qwen(in_ch, out_ch, base, z_dim, dim_mult, num_res_blocks, t_down):
    enc = _Holder()
    dims = [base * u for u in (1,) + tuple(dim_mult)]
    enc.conv_in = Conv(in_ch, dims[0], 3, kernel3d=(3, 3, 3))
    blocks = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            blocks.append(_qwen_resblock(cin, cout))
            cin = cout
        if i != len(dim_mult) - 1:
            blocks.append(_qwen_resample(cout, "downsample3d" if t_down[i] else "downsample2d"))
    enc.down_blocks = nn.ModuleList(blocks)
    enc.mid_block = _qwen_mid(dims[-1])
    enc.norm_out = RMSNorm(dims[-1], images=False)
    enc.conv_out = Conv(dims[-1], 2 * z_dim, 3, kernel3d=(3, 3, 3))

    dec = _Holder()
    ddims = [base * u for u in (dim_mult[-1],) + tuple(dim_mult[::-1])]
    dec.conv_in = Conv(z_dim, ddims[0], 3, kernel3d=(3, 3, 3))
    dec.mid_block = _qwen_mid(ddims[0])
    t_up = tuple(t_down[::-1])
    ups = []
    for i, (cin, cout) in enumerate(zip(ddims[:-1], ddims[1:])):
        if i > 0:
            cin = cin // 2
        b = _Holder()
        b.resnets = nn.ModuleList([_qwen_resblock(cin if j == 0 else cout, cout) for j in range(num_res_blocks + 1)])
        if i != len(dim_mult) - 1:
            b.upsamplers = nn.ModuleList([_qwen_resample(cout, "upsample3d" if t_up[i] else "upsample2d")])
        ups.append(b)
    dec.up_blocks = nn.ModuleList(ups)
    dec.norm_out = RMSNorm(ddims[-1], images=False)
    dec.conv_out = Conv(ddims[-1], out_ch, 3, kernel3d=(3, 3, 3))
    return enc, dec


# ------------------------------------------------------------------------------------------
# outputs / config
# ------------------------------------------------------------------------------------------
class _Stream:
    """The residual stream between blocks: the raw tensor and, when the producing conv could fuse it, the
    already normalised + activated copy for the consumer identified by ``act_key = (id(norm), silu)``."""
    __slots__ = ("raw", "act", "act_key", "gn_stats")

    def __init__(self, raw, act=None, act_key=None, gn_stats=None):
        self.raw, self.act, self.act_key = raw, act, act_key
        self.gn_stats = gn_stats  # (groups, [N, groups, 2] fp64 sums of ``raw``) when the producing conv's epilogue left them


class AutoencoderKLOutput(SimpleNamespace):
    """``.latent_dist`` like diffusers' output dataclass."""


class DecoderOutput(SimpleNamespace):
    """``.sample`` like diffusers' output dataclass."""


class VaeConfig(dict):
    """Attribute- and item-addressable, mutable (the reference assigns ``config.in_channels = 4``)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


FLUX_CONFIG = dict(
    _class_name="AutoencoderKL", in_channels=3, out_channels=3, latent_channels=16,
    block_out_channels=[128, 256, 512, 512], layers_per_block=2, act_fn="silu", norm_num_groups=32,
    down_block_types=["DownEncoderBlock2D"] * 4, up_block_types=["UpDecoderBlock2D"] * 4,
    sample_size=1024, scaling_factor=0.3611, shift_factor=0.1159, use_quant_conv=False, use_post_quant_conv=False,
    mid_block_add_attention=True, force_upcast=True)
QWEN_CONFIG = dict(
    _class_name="AutoencoderKLQwenImage", in_channels=3, out_channels=3, base_dim=96, z_dim=16, latent_channels=16,
    dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True], dropout=0.0,
    block_out_channels=[96, 192, 384, 384], sample_size=256, scaling_factor=1.0, shift_factor=0.0)


def _load_checkpoint_tensors(root: str) -> Dict[str, torch.Tensor]:
    """The weight files diffusers' ``from_pretrained`` accepts, in its order of preference: one safetensors file, a
    sharded safetensors checkpoint (``*.safetensors.index.json`` -> ``weight_map``), or a pickled state dict
    (``diffusion_pytorch_model.bin`` / ``pytorch_model.bin``, loaded with ``weights_only=True``)."""
    from safetensors.torch import load_file

    single = os.path.join(root, WEIGHTS_NAME)
    if os.path.isfile(single):
        return load_file(single)
    index = single + ".index.json"
    if os.path.isfile(index):
        with open(index) as f:
            shards = sorted(set(json.load(f)["weight_map"].values()))
        sd: Dict[str, torch.Tensor] = {}
        for shard in shards:
            sd.update(load_file(os.path.join(root, shard)))
        return sd
    for name in ("diffusion_pytorch_model.bin", "pytorch_model.bin"):
        cand = os.path.join(root, name)
        if os.path.isfile(cand):
            return dict(torch.load(cand, map_location="cpu", weights_only=True))
    raise FileNotFoundError(f"no {WEIGHTS_NAME}, sharded index or .bin state dict under {root}")


def _arch_of(config: dict) -> str:
    cls = str(config.get("_class_name", ""))
    if "Qwen" in cls or "base_dim" in config:
        return "qwen"
    return "flux"


# ------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------
class RgbaAutoencoder(nn.Module):
    def __init__(self, arch: str = "qwen", in_channels: int = 4, out_channels: int = 4, **overrides):
        super().__init__()
        if arch not in ("qwen", "flux"):
            raise ValueError(f"unknown arch {arch!r} (expected 'qwen' or 'flux')")
        self.arch = arch
        cfg = dict(FLUX_CONFIG if arch == "flux" else QWEN_CONFIG)
        cfg.update(overrides)
        cfg["in_channels"], cfg["out_channels"] = in_channels, out_channels
        self.config = VaeConfig(cfg)
        if arch == "flux":
            self.encoder, self.decoder = _build_flux(in_channels, out_channels, cfg["latent_channels"],
                                                     list(cfg["block_out_channels"]))
        else:
            z = cfg["z_dim"]
            self.encoder, self.decoder = _build_qwen(in_channels, out_channels, cfg["base_dim"], z, tuple(cfg["dim_mult"]),
                                                     cfg["num_res_blocks"], tuple(cfg["temperal_downsample"]))
            self.quant_conv = Conv(2 * z, 2 * z, 1, kernel3d=(1, 1, 1))
            self.post_quant_conv = Conv(z, z, 1, kernel3d=(1, 1, 1))
        self.use_tiling = False
        self.use_slicing = False
        self.gradient_checkpointing = False
        self.fuse_norm_residual = False
        self.fused_attention = True
        self.hpack_stem = True    # conv_in as 3 vertical taps over a horizontally packed 16-channel image (see _stem)
        self.fuse_norm = True  # RMS norm + SiLU in the producing conv's epilogue where one tile holds all channels
        self.fuse_gn_stats = True  # GroupNorm statistics out of the producing conv's epilogue (256 / 512-channel layers)
        self._pack_cache: Dict[tuple, tuple] = {}
        self.weights_generation = 0
        self.requires_grad_(False)
        self.eval()

    # ---- diffusers surface ---------------------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def enable_tiling(self, *a, **k):
        """diffusers tiling: inputs larger than the tile are processed in overlapping tiles whose seams are blended
        linearly (rgba_vae_stage.py:296-299 turns this on for training; results differ from the untiled path)."""
        self.use_tiling = True

    def _tiling(self):
        """(sample tile, sample stride, latent tile, latent stride) as the diffusers classes derive them."""
        if self.arch == "flux":   # AutoencoderKL: tile = sample_size, overlap factor 0.25
            ts = int(self.config.sample_size)
            tl = int(ts / (2 ** (len(self.config.block_out_channels) - 1)))
            return ts, int(ts * 0.75), tl, int(tl * 0.75)
        ts = int(getattr(self.config, "tile_sample_min_size", 256))   # AutoencoderKLQwenImage: 256 / stride 192
        st = int(getattr(self.config, "tile_sample_stride", 192))
        return ts, st, ts // 8, st // 8

    def _tiled(self, t: torch.Tensor, fn, tile: int, stride: int, blend: int, limit: int) -> torch.Tensor:
        """Overlapping tiles -> fn -> blend_v / blend_h in row-major order (in place, like diffusers) -> crop -> concat."""
        rows = []
        for i in range(0, t.shape[2], stride):
            row = []
            for j in range(0, t.shape[3], stride):
                row.append(fn(t[:, :, i:i + tile, j:j + tile].contiguous()))
            rows.append(row)
        out_rows = []
        for i, row in enumerate(rows):
            out_row = []
            for j, tl in enumerate(row):
                if i > 0:
                    ops.blend_tiles(rows[i - 1][j], tl, blend, vertical=True)
                if j > 0:
                    ops.blend_tiles(row[j - 1], tl, blend, vertical=False)
                out_row.append(tl[:, :, :limit, :limit])
            out_rows.append(torch.cat(out_row, dim=3))
        return torch.cat(out_rows, dim=2)

    def disable_tiling(self):
        self.use_tiling = False

    def enable_slicing(self):
        self.use_slicing = True  # numerically a no-op (per-sample split): honoured as a memory knob

    def disable_slicing(self):
        self.use_slicing = False

    def enable_gradient_checkpointing(self):
        """Honoured by ``trainer.VaeTrainStep``: residual and attention blocks keep only their input and recompute their
        intermediates in the backward (diffusers checkpoints per block; rgba_vae_stage.py:305-307 turns it on)."""
        self.gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        self.gradient_checkpointing = False

    def _encode_maybe_tiled(self, x: torch.Tensor, in_scale: float = 1.0, in_shift: float = 0.0) -> torch.Tensor:
        ts, ss, tl, sl = self._tiling()
        enc = lambda t: self._encode_moments(t, in_scale=in_scale, in_shift=in_shift)
        if self.use_tiling and isinstance(x, torch.Tensor) and x.dim() == 4 and (x.shape[-1] > ts or x.shape[-2] > ts):
            if self.arch == "flux":   # blend_extent = int(latent_tile * 0.25), row_limit = latent_tile - blend_extent
                blend = int(tl * 0.25)
                limit = tl - blend
            else:                      # blend = latent_tile - latent_stride, crop to the latent stride
                blend, limit = tl - sl, sl
            m = self._tiled(x, enc, ts, ss, blend, limit)
            return m[:, :, :x.shape[2] // 8, :x.shape[3] // 8].contiguous()
        return enc(x)

    def _decode_maybe_tiled(self, z: torch.Tensor, out_scale: float = 1.0, out_shift: float = 0.0, clamp=None) -> torch.Tensor:
        """``clamp((decode(z) * out_scale + out_shift))``.  Tiled: the affine map commutes with the linear seam blend, so
        it stays fused in every tile's conv_out epilogue; the clamp does not and runs once after blending."""
        ts, ss, tl, sl = self._tiling()
        if self.use_tiling and isinstance(z, torch.Tensor) and z.dim() == 4 and (z.shape[-1] > tl or z.shape[-2] > tl):
            if self.arch == "flux":
                blend = int(ts * 0.25)
                limit = ts - blend
                fn = lambda t: self._decode_image(t, out_scale=out_scale, out_shift=out_shift)
            else:
                blend, limit = ts - ss, ss
                # AutoencoderKLQwenImage.tiled_decode returns the blended tiles without the clamp of _decode
                fn = lambda t: self._decode_image(t, out_scale=out_scale, out_shift=out_shift, model_clamp=False)
            y = self._tiled(z, fn, tl, sl, blend, limit)
            y = y[:, :, :z.shape[2] * 8, :z.shape[3] * 8].contiguous()
            return y if clamp is None else y.clamp_(clamp[0], clamp[1])
        return self._decode_image(z, out_scale=out_scale, out_shift=out_shift, clamp=clamp)

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        moments = self._run_sliced(self._encode_maybe_tiled, x)
        post = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(latent_dist=post) if return_dict else (post,)

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        y = self._run_sliced(self._decode_maybe_tiled, z)
        return DecoderOutput(sample=y) if return_dict else (y,)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True, generator=None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator=generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)

    def _run_sliced(self, fn, t: torch.Tensor):
        if self.use_slicing and t.shape[0] > 1:
            return torch.cat([fn(s) for s in t.split(1)], dim=0)
        return fn(t)

    # ---- checkpoint I/O ------------------------------------------------------------------
    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **_):
        from safetensors.torch import save_file

        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: v for k, v in self.config.items()}
        # diffusers writes the config the class was REGISTERED with: a model widened by attribute
        # assignment still says 3.  We persist what the object says now AND keep loading tolerant.
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        save_file(sd, os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, torch_dtype: Optional[torch.dtype] = None,
                        ignore_mismatched_sizes: bool = False, low_cpu_mem_usage: bool = False, arch: Optional[str] = None,
                        **_):
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, CONFIG_NAME)) as f:
            cfg = json.load(f)
        arch = arch or _arch_of(cfg)
        in_ch = int(cfg.get("in_channels", cfg.get("input_channels", 3)))
        out_ch = int(cfg.get("out_channels", in_ch))
        keep = {k: v for k, v in cfg.items() if k not in ("in_channels", "out_channels")}
        model = cls(arch, in_ch, out_ch, **keep)
        sd = _load_checkpoint_tensors(root)
        own = model.state_dict()
        mismatched = [k for k, v in sd.items() if k in own and tuple(own[k].shape) != tuple(v.shape)]
        if mismatched and not ignore_mismatched_sizes:
            raise RuntimeError(f"size mismatch for {mismatched}; pass ignore_mismatched_sizes=True to keep the "
                               "newly initialised tensors (the reference then restores the RGBA convs itself)")
        if mismatched:
            warnings.warn(f"Some weights were not used because of a size mismatch and are newly initialised: {mismatched}")
        filtered = {k: v for k, v in sd.items() if k not in mismatched}
        missing, unexpected = model.load_state_dict(filtered, strict=False)
        missing = [k for k in missing if k not in mismatched]
        if missing or unexpected:
            warnings.warn(f"checkpoint/model key mismatch: missing={missing} unexpected={unexpected}")
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model

    # ---- weight packing cache ------------------------------------------------------------
    def mark_weights_changed(self) -> None:
        """Call after parameters were rewritten behind torch's back (raw-pointer optimizer updates: neither ``data_ptr``
        nor ``_version`` moves).  Drops every packed copy and bumps ``weights_generation`` (captured inference graphs of
        ``RgbaVAE.forward_graphed`` are keyed on it)."""
        self._pack_cache.clear()
        self.weights_generation += 1

    def _cached(self, key: tuple, params, build):
        ver = tuple((p.data_ptr(), p._version, tuple(p.shape), p.dtype) for p in params)
        hit = self._pack_cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = build()
        self._pack_cache[key] = (ver, val)
        return val

    def _f32(self, p: torch.Tensor, name: str) -> torch.Tensor:
        return self._cached((id(p), name), [p], lambda: p.detach().to(torch.float32).reshape(-1).contiguous())

    def _conv_weights(self, conv: Conv, tc: bool, upsample: bool = False, cin_pad: int = 0):
        def build():
            w = conv.weight2d().detach().to(torch.float32)
            if cin_pad and cin_pad > w.shape[1]:
                w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin_pad - w.shape[1]))
            w = w.contiguous()
            return ops.pack_conv_weights_tc(w, upsample) if tc else ops.pack_conv_weights_direct(w)

        return self._cached((id(conv), "tc" if tc else "direct", upsample, cin_pad), [conv.weight], build)

    def _hpack_weights(self, conv: Conv):
        """[cout][c][3][3] -> bf16 [cout][3 * 16]: per vertical tap dy the 16-wide row [dx][c] (zero padded) that matches
        ``ops.nchw_to_nhwc_hpack``."""
        def build():
            w = conv.weight2d().detach().to(torch.float32)            # [cout][c][dy][dx]
            m = w.permute(0, 2, 3, 1).reshape(w.shape[0], 3, -1)      # [cout][dy][dx*c + ch]
            m = torch.nn.functional.pad(m, (0, 16 - m.shape[2]))
            return m.reshape(w.shape[0], 48).to(torch.bfloat16).contiguous()

        return self._cached((id(conv), "hpack"), [conv.weight], build)

    # ---- building blocks -----------------------------------------------------------------
    def _mode(self):
        dt = self.dtype
        if dt == torch.bfloat16:
            return True, torch.bfloat16, RV_BF16
        if dt == torch.float32:
            return False, torch.float32, RV_F32
        raise TypeError(f"RgbaAutoencoder runs in float32 or bfloat16, not {dt}")

    def _conv(self, x: torch.Tensor, conv: Conv, *, upsample: bool = False, residual: Optional[torch.Tensor] = None,
              y_nchw: bool = False, y_dtype: Optional[torch.dtype] = None, out_scale: float = 1.0, out_shift: float = 0.0,
              clamp=None) -> torch.Tensor:
        """x: NHWC [N,H,W,Cin(+pad)] activations in the model dtype."""
        return self._conv_fused(x, conv, upsample=upsample, residual=residual, y_nchw=y_nchw, y_dtype=y_dtype,
                                out_scale=out_scale, out_shift=out_shift, clamp=clamp).raw

    def _can_fuse_norm(self, conv: Conv, norm) -> bool:
        """The conv epilogue can apply the consumer's norm when it is a per-pixel RMS norm and one accumulator
        tile holds the pixel's whole channel vector (tensor-core path, Cout <= 256)."""
        return self.fuse_norm and self.dtype == torch.bfloat16 and isinstance(norm, RMSNorm) and conv.out_channels <= 256

    def _conv_fused(self, x: torch.Tensor, conv: Conv, *, upsample: bool = False, residual: Optional[torch.Tensor] = None,
                    y_nchw: bool = False, y_dtype: Optional[torch.dtype] = None, out_scale: float = 1.0,
                    out_shift: float = 0.0, clamp=None, next_norm=None, want_raw: bool = True,
                    hpack: bool = False) -> "_Stream":
        """Convolution whose result feeds ``next_norm = (norm_module, silu)``: where possible that norm is fused
        into the epilogue (second output ``act``); ``want_raw=False`` drops the raw tensor when only the
        normalised one is consumed (conv1 -> norm2 -> conv2 inside a residual block)."""
        tc, act_dt, code = self._mode()
        n, h, w, cx = x.shape
        cin, cout, k, stride = conv.in_channels, conv.out_channels, conv.k, conv.stride
        oh, ow = ops.conv_out_size(h, w, k, stride, upsample)
        y_dt = act_dt if y_dtype is None else y_dtype
        bias = self._f32(conv.bias, "bias")
        desc = ops.make_desc(n, h, w, cx if tc else cin, cout, k, stride, upsample, x_dtype=code,
                             y_dtype=RV_F32 if y_dt == torch.float32 else RV_BF16, y_nchw=y_nchw, x_cstride=cx,
                             out_scale=out_scale, out_shift=out_shift, clamp=clamp, taps_1d=hpack)
        fuse = next_norm is not None and not y_nchw and y_dt == torch.bfloat16 and self._can_fuse_norm(conv, next_norm[0])
        if fuse and residual is not None and not self.fuse_norm_residual:
            # measured: residual add + raw store + two-pass norm in one epilogue is slower than conv + norm kernel
            fuse = False
        y = None
        if want_raw or not fuse:
            y = torch.empty((n, cout, oh, ow) if y_nchw else (n, oh, ow, cout), dtype=y_dt, device=x.device)
        if tc and y_nchw and residual is None and not hpack and ops.conv_out_eligible(n, h, w, cin, cx, cout, k, stride, upsample):
            # conv_out: an HBM-bound layer (Cout = 4) on its own kernel -- the kernel-row index rides in the MMA's N
            wt = self._cached((id(conv), "conv_out_taps"), [conv.weight], lambda: ops.pack_conv_out_weights(conv.weight2d()))
            ops.conv_out(desc, x, wt, bias, y)
            return _Stream(y, None, None)
        if tc:
            if hpack:
                wp = self._hpack_weights(conv)
            else:
                # stride-2 convs whose channel count is not a multiple of 64 (Qwen's 96 -> 96 down-sampler): every tap's K range
                # padded to 128 with zero weights, so that the kernel can use 128-byte operand rows on the parity view
                pad_k = (cx + 63) // 64 * 64 if (stride == 2 and k == 3 and cx > 64 and cx % 64) else cx
                wp = self._conv_weights(conv, True, upsample, cin_pad=pad_k)
            if fuse:
                norm, silu = next_norm
                act = torch.empty((n, oh, ow, cout), dtype=y_dt, device=x.device)
                gamma = self._cached((id(norm.gamma), "gamma_scaled"), [norm.gamma],
                                     lambda: (norm.gamma.detach().to(torch.float32).reshape(-1) * math.sqrt(cout)).contiguous())
                ops.conv2d_tc_norm(desc, x, wp, wp.shape[1], bias, residual, y, act, gamma, silu)
                return _Stream(y, act, (id(norm), silu))
            if (self.fuse_gn_stats and next_norm is not None and isinstance(next_norm[0], GroupNorm) and not y_nchw
                    and y_dt == torch.bfloat16 and not hpack):
                # the consumer is a GroupNorm: its statistics come out of this conv's epilogue where the layer has that form
                groups = int(next_norm[0].num_groups)
                stats = ops.conv2d_tc_gnstats(desc, x, wp, wp.shape[1], bias, residual, y, groups)
                if stats is not None:
                    return _Stream(y, None, None, (groups, stats))
            ops.conv2d_tc(desc, x, wp, wp.shape[1], bias, residual, y)
        else:
            wp = self._conv_weights(conv, False)
            ops.conv2d_direct(desc, x, wp, bias, residual, y)
        return _Stream(y, None, None)

    def _normed(self, st: "_Stream", norm, silu: bool = True) -> torch.Tensor:
        """act(norm(stream)): the fused copy if the producer already wrote it for this norm, else the norm kernel."""
        if st.act is not None and st.act_key == (id(norm), silu):
            return st.act
        stats = None
        if st.gn_stats is not None and not isinstance(norm, RMSNorm) and st.gn_stats[0] == norm.num_groups:
            stats = st.gn_stats[1]
        return self._norm(st.raw, norm, silu, stats=stats)

    def _norm(self, x: torch.Tensor, norm, silu: bool = True, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        if isinstance(norm, RMSNorm):
            return ops.rmsnorm_silu(x, self._f32(norm.gamma, "gamma"), silu)
        return ops.groupnorm_silu(x, self._f32(norm.weight, "gn_w"), self._f32(norm.bias, "gn_b"), norm.num_groups, norm.eps,
                                  silu, stats=stats)

    def _resblock(self, st: "_Stream", blk, next_norm=None) -> "_Stream":
        """h = shortcut(x); x = conv2(act(norm2(conv1(act(norm1(x)))))) + h   (diffusers ResnetBlock2D /
        QwenImageResidualBlock).  conv1 writes only act(norm2(.)), conv2 adds h and also writes the next consumer's norm."""
        short = getattr(blk, "conv_shortcut", None)
        a = self._normed(st, blk.norm1)
        h = st.raw if short is None else self._conv(st.raw, short)
        t = self._conv_fused(a, blk.conv1, next_norm=(blk.norm2, True), want_raw=False)
        return self._conv_fused(self._normed(t, blk.norm2), blk.conv2, residual=h, next_norm=next_norm)

    def _gemm(self, x, w, *, rows, k, cols, x_ld, w_ld, y, y_ld, bias=None, bias_mode=0, alpha=1.0, residual=None):
        """y[rows][cols] = alpha * x[rows][k] . w[cols][k]^T (+bias) on the conv kernels (1x1, one 'image row')."""
        tc, _, code = self._mode()
        desc = ops.make_desc(1, 1, rows, k, cols, 1, 1, False, x_dtype=code,
                             y_dtype=RV_F32 if y.dtype == torch.float32 else RV_BF16, x_cstride=x_ld, y_cstride=y_ld,
                             bias_mode=bias_mode, alpha=alpha)
        if tc:
            ops.conv2d_tc(desc, x, w, w_ld, bias, residual, y)
        else:
            assert w_ld == k, "direct path needs dense weight rows"
            ops.conv2d_direct(desc, x, w, bias, residual, y)

    def _attention(self, st: "_Stream", attn) -> "_Stream":
        """Single-head self-attention over the H*W tokens of each image, d = C, residual added."""
        tc, act_dt, _ = self._mode()
        x = st.raw
        n, h, w, c = x.shape
        t = h * w
        dev = x.device
        if self.arch == "qwen":
            xn = self._normed(st, attn.norm, silu=False)
            wqkv = attn.to_qkv.weight.detach().reshape(3 * c, c)
            bqkv = attn.to_qkv.bias.detach()
            srcs = [attn.to_qkv.weight, attn.to_qkv.bias]
            get = lambda i: (wqkv[i * c:(i + 1) * c], bqkv[i * c:(i + 1) * c])
            wo_p, bo_p = attn.proj.weight, attn.proj.bias
        else:
            xn = self._normed(st, attn.group_norm, silu=False)
            lins = (attn.to_q, attn.to_k, attn.to_v)
            srcs = [p for l in lins for p in (l.weight, l.bias)]
            get = lambda i: (lins[i].weight.detach(), lins[i].bias.detach())
            wo_p, bo_p = attn.to_out[0].weight, attn.to_out[0].bias

        def build():
            f = lambda a: a.to(torch.float32).contiguous()
            (wq, bq), (wk, bk), (wv, bv) = get(0), get(1), get(2)
            wt = torch.bfloat16 if tc else torch.float32
            return dict(wqk=torch.cat([f(wq), f(wk)], 0).to(wt).contiguous(), bqk=torch.cat([f(bq), f(bk)], 0).contiguous(),
                        wq=f(wq).to(wt), wk=f(wk).to(wt), bq=f(bq), bk=f(bk), wv=f(wv).to(wt).contiguous(), bv=f(bv),
                        wo=f(wo_p.detach().reshape(c, c)).to(wt).contiguous(), bo=f(bo_p.detach()))

        pk = self._cached((id(attn), "attn", tc), srcs + [wo_p, bo_p], build)
        scale = ops.attn_scale(c)
        xn2 = xn.view(n * t, c)
        o = torch.empty((n * t, c), dtype=act_dt, device=dev)
        if tc:
            qk = torch.empty((n * t, 2 * c), dtype=act_dt, device=dev)
            self._gemm(xn2, pk["wqk"], rows=n * t, k=c, cols=2 * c, x_ld=c, w_ld=c, y=qk, y_ld=2 * c, bias=pk["bqk"], bias_mode=1)
            q_all, k_all, ld_qk = qk, qk[:, c:], 2 * c
        else:
            q_all = torch.empty((n * t, c), dtype=act_dt, device=dev)
            k_all = torch.empty((n * t, c), dtype=act_dt, device=dev)
            self._gemm(xn2, pk["wq"], rows=n * t, k=c, cols=c, x_ld=c, w_ld=c, y=q_all, y_ld=c, bias=pk["bq"], bias_mode=1)
            self._gemm(xn2, pk["wk"], rows=n * t, k=c, cols=c, x_ld=c, w_ld=c, y=k_all, y_ld=c, bias=pk["bk"], bias_mode=1)
            ld_qk = c
        if tc and self.fused_attention and c in ops.FUSED_ATTENTION_DIMS and t % 128 == 0:
            # flash-style kernel: scores / probabilities never reach HBM
            # V^T for ALL images from one GEMM: [c][n * t] (per-image launches were 8 x 25 us of latency per attention block)
            vt_all = torch.empty((c, n * t), dtype=act_dt, device=dev)
            self._gemm(pk["wv"], xn2, rows=c, k=c, cols=n * t, x_ld=c, w_ld=c, y=vt_all, y_ld=n * t, bias=pk["bv"], bias_mode=2)
            o = ops.attention(q_all[:, :c], k_all[:, :c], vt_all, n, t)
        else:
            self._attention_unfused(pk, xn2, q_all, k_all, ld_qk, o, n, t, c, scale, act_dt, dev)
        out = torch.empty_like(x)
        self._gemm(o, pk["wo"], rows=n * t, k=c, cols=c, x_ld=c, w_ld=c, y=out.view(n * t, c), y_ld=c, bias=pk["bo"],
                   bias_mode=1, residual=x.view(n * t, c))
        return _Stream(out, None, None)

    def _attention_unfused(self, pk, xn2, q_all, k_all, ld_qk, o, n, t, c, scale, act_dt, dev):
        """QK^T GEMM -> fp32 scores -> softmax kernel -> PV GEMM (fp32 mode, d != 384, ragged token counts)."""
        vt = torch.empty((c, t), dtype=act_dt, device=dev)
        # bound the fp32 score block to ~1 GiB; chunks are multiples of 128 query rows
        q_chunk = max(128, min(t, ((1 << 28) // t) // 128 * 128))
        s = torch.empty((min(q_chunk, t), t), dtype=torch.float32, device=dev)
        for i in range(n):
            xi = xn2[i * t:(i + 1) * t]
            # V^T[c][token] = Wv . xn^T + bv  (bias per GEMM row)
            self._gemm(pk["wv"], xi, rows=c, k=c, cols=t, x_ld=c, w_ld=c, y=vt, y_ld=t, bias=pk["bv"], bias_mode=2)
            for r0 in range(0, t, q_chunk):
                rows = min(q_chunk, t - r0)
                qi = q_all[i * t + r0:i * t + r0 + rows]
                ki = k_all[i * t:(i + 1) * t]
                sv = s[:rows]
                self._gemm(qi, ki, rows=rows, k=c, cols=t, x_ld=ld_qk, w_ld=ld_qk, y=sv, y_ld=t, alpha=scale)
                p = ops.softmax_rows(sv, act_dt)
                self._gemm(p, vt, rows=rows, k=t, cols=c, x_ld=t, w_ld=t, y=o[i * t + r0:i * t + r0 + rows], y_ld=c)

    def _mid(self, st, mid, next_norm=None):
        a = mid.attentions[0]
        st = self._resblock(st, mid.resnets[0], next_norm=(getattr(a, "norm", None) or a.group_norm, False))
        st = self._attention(st, a)
        return self._resblock(st, mid.resnets[1], next_norm=next_norm)

    # ---- encode / decode -----------------------------------------------------------------
    def _check_image(self, x: torch.Tensor, channels: int, what: str, mult: int):
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError(f"{what} must be a 4-D (B,C,H,W) tensor, got {tuple(getattr(x, 'shape', ()))}")
        if x.shape[1] != channels:
            raise ValueError(f"{what} has {x.shape[1]} channels, the model expects {channels}")
        if x.shape[2] % mult or x.shape[3] % mult:
            raise ValueError(f"{what} height and width must be multiples of {mult}, got {tuple(x.shape[2:])}")
        if not x.is_cuda:
            raise RvError("RgbaAutoencoder runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"{what} must be float32 or bfloat16, got {x.dtype}")

    def _stem(self, x: torch.Tensor, conv: Conv, in_scale: float, in_shift: float, next_norm=None) -> "_Stream":
        """NCHW boundary tensor -> first NHWC activation."""
        tc, act_dt, code = self._mode()
        n, c, h, w = x.shape
        x = x.contiguous()
        if tc and self.hpack_stem and conv.k == 3 and conv.stride == 1 and 3 * c <= 16:
            # few-channel 3x3 stem (conv_in, Cin = 4): the three horizontal neighbours ride in the pixel's 16 channels,
            # the conv runs as 3 vertical taps of K = 16 (3 MMAs + 3 TMA boxes per tile instead of 9 + 9)
            xp = ops.nchw_to_nhwc_hpack(x, in_scale, in_shift)
            return self._conv_fused(xp, conv, next_norm=next_norm, hpack=True)
        if tc:
            xp = ops.nchw_to_nhwc(x, 16 * ((c + 15) // 16), act_dt, in_scale, in_shift)
            return self._conv_fused(xp, conv, next_norm=next_norm)
        y = torch.empty((n, h, w, conv.out_channels), dtype=act_dt, device=x.device)
        desc = ops.make_desc(n, h, w, c, conv.out_channels, conv.k, 1, False,
                             x_dtype=RV_F32 if x.dtype == torch.float32 else RV_BF16, y_dtype=code, x_nchw=True,
                             in_scale=in_scale, in_shift=in_shift)
        ops.conv2d_direct(desc, x, self._conv_weights(conv, False), self._f32(conv.bias, "bias"), None, y)
        return _Stream(y, None, None)

    @staticmethod
    def _first_norm(item):
        """(norm, silu) the op ``item = (kind, module)`` applies to the stream first, or None."""
        kind, m = item
        if kind == "res":
            return (m.norm1, True)
        if kind == "mid":
            return (m.resnets[0].norm1, True)
        if kind == "head":
            return (m[0], True)
        return None

    def _encode_moments(self, x: torch.Tensor, in_scale: float = 1.0, in_shift: float = 0.0) -> torch.Tensor:
        """(B,Cin,H,W) in [-1,1] (after in_scale/in_shift) -> moments (B,2Z,H/8,W/8), model dtype."""
        self._check_image(x, self.encoder.conv_in.in_channels, "encode() input", 8)
        enc = self.encoder
        items = []
        if self.arch == "flux":
            for blk in enc.down_blocks:
                items += [("res", r) for r in blk.resnets]
                if getattr(blk, "downsamplers", None) is not None:
                    items.append(("down", blk.downsamplers[0].conv))
            items.append(("mid", enc.mid_block))
            head = (enc.conv_norm_out, enc.conv_out)
        else:
            for blk in enc.down_blocks:
                items.append(("res", blk) if blk._kind == "res" else ("down", blk.resample[1]))
            items.append(("mid", enc.mid_block))
            head = (enc.norm_out, enc.conv_out)
        items.append(("head", head))
        st = self._stem(x, enc.conv_in, in_scale, in_shift, next_norm=self._first_norm(items[0]))
        st = self._run_ops_until_head(st, items)
        h = self._normed(st, head[0])
        if self.arch == "flux":
            return self._conv(h, head[1], y_nchw=True)
        return self._conv(self._conv(h, head[1]), self.quant_conv, y_nchw=True)

    def _run_ops_until_head(self, st, items):
        """items[-1] is the ("head", (norm, conv)) marker: run everything before it with lookahead into it."""
        for i, (kind, m) in enumerate(items[:-1]):
            nxt = self._first_norm(items[i + 1])
            st = self._run_ops_one(st, kind, m, nxt)
        return st

    def _run_ops_one(self, st, kind, m, nxt):
        if kind == "res":
            return self._resblock(st, m, next_norm=nxt)
        if kind == "mid":
            return self._mid(st, m, next_norm=nxt)
        if kind == "down" or kind == "conv":
            return self._conv_fused(st.raw, m, next_norm=nxt)
        if kind == "up":
            return self._conv_fused(st.raw, m, upsample=True, next_norm=nxt)
        raise AssertionError(kind)

    def _decode_image(self, z: torch.Tensor, out_scale: float = 1.0, out_shift: float = 0.0, clamp=None,
                      z_scale: float = 1.0, z_shift: float = 0.0, model_clamp: bool = True) -> torch.Tensor:
        """latents (B,Z,h,w) -> image (B,Cout,8h,8w) in the model dtype (optionally y*out_scale+out_shift, clamped)."""
        dec = self.decoder
        self._check_image(z, dec.conv_in.in_channels, "decode() input", 1)
        items = [("mid", dec.mid_block)]
        for blk in dec.up_blocks:
            items += [("res", r) for r in blk.resnets]
            if getattr(blk, "upsamplers", None) is not None:
                up = blk.upsamplers[0]
                items.append(("up", up.conv if self.arch == "flux" else up.resample[1]))
        if self.arch == "flux":
            head = (dec.conv_norm_out, dec.conv_out)
            items.append(("head", head))
            st = self._stem(z, dec.conv_in, z_scale, z_shift, next_norm=self._first_norm(items[0]))
            st = self._run_ops_until_head(st, items)
            return self._conv(self._normed(st, head[0]), head[1], y_nchw=True, out_scale=out_scale, out_shift=out_shift,
                              clamp=clamp)
        head = (dec.norm_out, dec.conv_out)
        items.append(("head", head))
        st = self._stem(z, self.post_quant_conv, z_scale, z_shift)
        st = self._conv_fused(st.raw, dec.conv_in, next_norm=self._first_norm(items[0]))
        st = self._run_ops_until_head(st, items)
        # AutoencoderKLQwenImage._decode clamps to [-1, 1]; composed with the caller's affine + clamp
        if model_clamp:
            lo, hi = -1.0 * out_scale + out_shift, 1.0 * out_scale + out_shift
            if clamp is not None:
                lo, hi = max(lo, clamp[0]), min(hi, clamp[1])
            clamp = (lo, hi)
        return self._conv(self._normed(st, head[0]), head[1], y_nchw=True, out_scale=out_scale, out_shift=out_shift,
                          clamp=clamp)
