"""Tensor-level wrappers over the C ABI.  torch supplies device memory and streams only; every
function below launches librgbavae kernels on ``torch.cuda.current_stream()`` and raises if the
tensors are not CUDA tensors -- there is no eager path."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, RV_BF16, RV_F32, check


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return RV_F32
    if t.dtype == torch.bfloat16:
        return RV_BF16
    raise TypeError(f"librgbavae supports float32 and bfloat16 tensors, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.RvError("ragb_vae_b200 runs on CUDA (sm_100a) tensors only; there is no CPU path")


def launch_count() -> int:
    return int(_lib.load().rv_launch_count())


def prof_begin() -> None:
    check(_lib.load().rv_prof_begin(), "rv_prof_begin")


def prof_end() -> dict:
    n = _lib.PROF_CATEGORIES
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    work = (C.c_double * n)()
    check(_lib.load().rv_prof_end(ms, cnt, work), "rv_prof_end")
    return {name: {"ms": ms[i], "launches": int(cnt[i]), "work": work[i]} for i, name in enumerate(_lib.PROF_NAMES)}


# ------------------------------------------------------------------------------------------
# convolution / GEMM
# ------------------------------------------------------------------------------------------
def conv_out_size(h: int, w: int, ksize: int, stride: int, upsample: bool):
    if upsample:
        return 2 * h, 2 * w
    if ksize == 3 and stride == 2:
        return h // 2, w // 2
    return h, w


def make_desc(n, h, w, cin, cout, ksize=3, stride=1, upsample=False, *, x_dtype, y_dtype, x_nchw=False, y_nchw=False,
              x_cstride=None, y_cstride=None, bias_mode=1, in_scale=1.0, in_shift=0.0, out_scale=1.0, out_shift=0.0,
              clamp=None, alpha=1.0, taps_1d=False) -> ConvDesc:
    oh, ow = conv_out_size(h, w, ksize, stride, upsample)
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout = n, h, w, cin, cout
    d.ksize, d.stride, d.upsample = ksize, stride, int(bool(upsample))
    d.pad_lo = 1 if (ksize == 3 and stride == 1) else 0
    d.oh, d.ow = oh, ow
    d.x_dtype, d.y_dtype = x_dtype, y_dtype
    d.x_nchw, d.y_nchw = int(x_nchw), int(y_nchw)
    d.x_cstride = cin if x_cstride is None else x_cstride
    d.y_cstride = cout if y_cstride is None else y_cstride
    d.bias_mode = bias_mode
    d.in_scale, d.in_shift, d.out_scale, d.out_shift = in_scale, in_shift, out_scale, out_shift
    d.clamp = 0 if clamp is None else 1
    d.clamp_lo, d.clamp_hi = (0.0, 0.0) if clamp is None else clamp
    d.alpha = alpha
    d.taps_1d = int(bool(taps_1d))
    return d


def conv2d_direct(desc: ConvDesc, x, w, bias, residual, y) -> None:
    _need_cuda(x, w, bias, residual, y)
    check(_lib.load().rv_conv2d_direct(C.byref(desc), _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), _ptr(y), _stream(x)),
          "rv_conv2d_direct")


def conv2d_tc(desc: ConvDesc, x, w_packed, w_ld: int, bias, residual, y) -> None:
    _need_cuda(x, w_packed, bias, residual, y)
    check(_lib.load().rv_conv2d_tc(C.byref(desc), _ptr(x), _ptr(w_packed), int(w_ld), _ptr(bias), _ptr(residual), _ptr(y),
                                   _stream(x)), "rv_conv2d_tc")


_GN_SCRATCH: dict = {}


def conv2d_tc_gnstats(desc: ConvDesc, x, w_packed, w_ld: int, bias, residual, y, groups: int = 32):
    """conv2d_tc whose epilogue also leaves the GroupNorm statistics of y: returns the [N, groups, 2] fp64 (sum, sum of
    squares) tensor ``groupnorm_silu(..., stats=)`` takes, or None where the layer has no statistics epilogue (nothing is
    launched then: the caller runs ``conv2d_tc``)."""
    _need_cuda(x, w_packed, bias, residual, y)
    lib = _lib.load()
    need = int(lib.rv_conv2d_tc_gnstats_scratch_bytes(C.byref(desc), int(groups)))
    if need == 0 or bias is None:
        return None
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
    scratch = _GN_SCRATCH.get(key)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty((need,), dtype=torch.uint8, device=x.device)
        _GN_SCRATCH[key] = scratch
    stats = torch.empty((desc.n, groups, 2), dtype=torch.float64, device=x.device)
    check(lib.rv_conv2d_tc_gnstats(C.byref(desc), _ptr(x), _ptr(w_packed), int(w_ld), _ptr(bias), _ptr(residual), _ptr(y),
                                   int(groups), _ptr(stats), _ptr(scratch), scratch.numel(), _stream(x)), "rv_conv2d_tc_gnstats")
    return stats


def conv_out(desc: ConvDesc, x, w_taps, bias, y) -> None:
    """The decoder's conv_out (3x3, 64/96/128 -> <= 5 channels, NCHW output) on its HBM-bound kernel (rv_conv_out)."""
    _need_cuda(x, w_taps, bias, y)
    check(_lib.load().rv_conv_out(C.byref(desc), _ptr(x), _ptr(w_taps), _ptr(bias), _ptr(y), _stream(x)), "rv_conv_out")


def pack_conv_out_weights(w: torch.Tensor) -> torch.Tensor:
    """[cout][cin][3][3] -> bf16 [3 (dx)][16 rows: dy * cout + co, zero padded][cin] as one [48][cin] matrix."""
    cout, cin = w.shape[0], w.shape[1]
    if 3 * cout > 16 or tuple(w.shape[2:]) != (3, 3):
        raise ValueError(f"conv_out weights must be [<=5][cin][3][3], got {tuple(w.shape)}")
    t = w.detach().to(torch.float32).permute(3, 2, 0, 1).reshape(3, 3 * cout, cin)   # [dx][dy * cout + co][c]
    t = torch.nn.functional.pad(t, (0, 0, 0, 16 - 3 * cout))
    return t.reshape(48, cin).to(torch.bfloat16).contiguous()


def conv_out_eligible(n: int, h: int, w: int, cin: int, cx: int, cout: int, k: int, stride: int, upsample: bool) -> bool:
    return k == 3 and stride == 1 and not upsample and cin in (64, 96, 128) and cx == cin and 3 * cout <= 16 and w >= 64


def conv2d_tc_norm(desc: ConvDesc, x, w_packed, w_ld: int, bias, residual, y, y_act, gamma_scaled, silu: bool) -> None:
    """conv2d_tc with the consumer's RMS norm (+SiLU) fused into the epilogue; y may be None (activated output only)."""
    _need_cuda(x, w_packed, bias, residual, y, y_act, gamma_scaled)
    check(_lib.load().rv_conv2d_tc_norm(C.byref(desc), _ptr(x), _ptr(w_packed), int(w_ld), _ptr(bias), _ptr(residual),
                                        _ptr(y), _ptr(y_act), _ptr(gamma_scaled), int(silu), _stream(x)),
          "rv_conv2d_tc_norm")


def pack_conv_weights_tc(w: torch.Tensor, upsample: bool = False) -> torch.Tensor:
    """fp32 [cout][cin][k][k] -> bf16 [cout][taps*cin] (upsample: [cout][16*cin], phase-folded)."""
    _need_cuda(w)
    w = w.detach().to(torch.float32).contiguous()
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    slots = 16 if upsample else k * k
    out = torch.empty((cout, slots * cin), dtype=torch.bfloat16, device=w.device)
    ld = C.c_int64(0)
    check(_lib.load().rv_pack_conv_weights(_ptr(w), cout, cin, k, int(upsample), _ptr(out), C.byref(ld), _stream(w)),
          "rv_pack_conv_weights")
    assert ld.value == slots * cin
    return out


def pack_conv_weights_direct(w: torch.Tensor) -> torch.Tensor:
    """fp32 [cout][cin][k][k] -> fp32 [cout][taps*cin]."""
    _need_cuda(w)
    w = w.detach().to(torch.float32).contiguous()
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    out = torch.empty((cout, k * k * cin), dtype=torch.float32, device=w.device)
    check(_lib.load().rv_pack_conv_weights_direct(_ptr(w), cout, cin, k, _ptr(out), _stream(w)),
          "rv_pack_conv_weights_direct")
    return out


# ------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------
def rmsnorm_silu(x: torch.Tensor, gamma: torch.Tensor, silu: bool = True, out: Optional[torch.Tensor] = None):
    """x: NHWC-dense [..., C]."""
    _need_cuda(x, gamma)
    c = x.shape[-1]
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().rv_rmsnorm_silu(_ptr(x), _ptr(gamma), _ptr(y), x.numel() // c, c, _dt(x), int(silu), _stream(x)),
          "rv_rmsnorm_silu")
    return y


def groupnorm_silu(x: torch.Tensor, gamma, beta, groups: int = 32, eps: float = 1e-6, silu: bool = True,
                   out: Optional[torch.Tensor] = None, return_stats: bool = False, stats: Optional[torch.Tensor] = None):
    """x: [N, H, W, C] (or [N, HW, C]) NHWC-dense.  ``return_stats``: also the [N, groups, 2] fp64 (sum, sum of squares)
    the backward pass needs.  ``stats``: those sums if the producer of x already has them (``conv2d_tc_gnstats``): the
    statistics pass over x is skipped."""
    _need_cuda(x, gamma, beta, stats)
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    lib = _lib.load()
    if stats is None:
        stats = torch.empty((n, groups, 2), dtype=torch.float64, device=x.device)
        check(lib.rv_groupnorm_stats(_ptr(x), _ptr(stats), n, hw, c, groups, _dt(x), _stream(x)), "rv_groupnorm_stats")
    elif tuple(stats.shape) != (n, groups, 2) or stats.dtype != torch.float64 or not stats.is_contiguous():
        raise ValueError(f"groupnorm_silu: stats must be a contiguous float64 [{n}, {groups}, 2] tensor")
    y = torch.empty_like(x) if out is None else out
    check(lib.rv_groupnorm_silu(_ptr(x), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(y), n, hw, c, groups, eps, _dt(x),
                                int(silu), _stream(x)), "rv_groupnorm_silu")
    return (y, stats) if return_stats else y


FUSED_ATTENTION_D = 384            # Qwen-Image mid block
FUSED_ATTENTION_DIMS = (384, 512)  # + the Flux AutoencoderKL mid block (two output passes inside the kernel)


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, n_img: int, tokens: int, return_lse: bool = False):
    """Fused softmax(Q K^T / sqrt(d)) V.  q, k: bf16 [n_img*tokens, d] row views (may be column slices of one
    tensor, same row stride); vt: bf16 [n_img, d, tokens] or [d, n_img*tokens].  Returns bf16 [n_img*tokens, d]; with ``return_lse`` also the
    per-row base-2 log-sum-exp of the scaled scores (fp32 [n_img*tokens]) the attention backward recomputes P from."""
    _need_cuda(q, k, vt)
    # vt: dense [n_img, d, tokens], or [d, n_img * tokens] (what one projection GEMM over all images writes)
    if vt.dim() == 3:
        d, ld_vt, img_pitch = vt.shape[1], tokens, vt.shape[1] * tokens
        ok = vt.is_contiguous() and tuple(vt.shape) == (n_img, d, tokens)
    else:
        d, ld_vt, img_pitch = vt.shape[0], vt.stride(0), tokens
        ok = vt.dim() == 2 and vt.stride(1) == 1 and vt.shape[1] == n_img * tokens
    if q.stride(0) != k.stride(0) or q.stride(1) != 1 or k.stride(1) != 1 or not ok:
        raise ValueError("attention: q/k must share a row stride and be unit-stride in d; vt must be [n, d, t] dense or [d, n*t]")
    out = torch.empty((n_img * tokens, d), dtype=torch.bfloat16, device=q.device)
    lse = torch.empty((n_img * tokens,), dtype=torch.float32, device=q.device) if return_lse else None
    ws = _attention_workspace(q.device, tokens, d)
    check(_lib.load().rv_attention_ws(_ptr(q), _ptr(k), q.stride(0), _ptr(vt), ld_vt, img_pitch, _ptr(out), d, _ptr(lse), _ptr(ws),
                                      0 if ws is None else ws.numel(), n_img, tokens, d, _stream(q)), "rv_attention_ws")
    return (out, lse) if return_lse else out


_ATTN_WS: dict = {}


def _attention_workspace(device: torch.device, tokens: int, d: int) -> Optional[torch.Tensor]:
    """Per (device, stream) scratch of the d = 512 kernel (probability tiles of pass 1, replayed by pass 2); grow-only.
    Its first 1024 bytes are the slot flags: zeroed once here, left zero by every launch."""
    need = int(_lib.load().rv_attention_workspace_bytes(int(tokens), int(d)))
    if need == 0:
        return None
    key = (device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty((need,), dtype=torch.uint8, device=device)
        ws[:1024].zero_()
        _ATTN_WS[key] = ws
    return ws


def gemm_rowstat(x: torch.Tensor, w: torch.Tensor, rowstat: torch.Tensor, mode: int, alpha: float,
                 mul_in: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 y[rows][cols] from acc = x[rows][k] . w[cols][k]^T with the attention backward's elementwise step in the
    tensor-core epilogue: mode 1 -> exp2(alpha * acc - rowstat[row]); mode 2 -> mul_in * (alpha * acc - rowstat[row])."""
    _need_cuda(x, w, rowstat, mul_in, out)
    rows, k = x.shape
    cols = w.shape[0]
    if x.stride(1) != 1 or w.stride(1) != 1 or rowstat.dtype != torch.float32 or rowstat.numel() != rows or not rowstat.is_contiguous():
        raise ValueError("gemm_rowstat: unit-stride bf16 operands and one fp32 statistic per row")
    y = torch.empty((rows, cols), dtype=torch.bfloat16, device=x.device) if out is None else out
    if mul_in is not None and (mul_in.shape != y.shape or mul_in.stride(0) != y.stride(0) or mul_in.dtype != torch.bfloat16):
        raise ValueError("gemm_rowstat: mul_in must have the shape, pitch and dtype of the output")
    desc = make_desc(1, 1, rows, k, cols, 1, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16, x_cstride=x.stride(0),
                     y_cstride=y.stride(0), bias_mode=0, alpha=alpha)
    check(_lib.load().rv_gemm_rowstat(C.byref(desc), _ptr(x), _ptr(w), w.stride(0), _ptr(rowstat), int(mode), _ptr(mul_in),
                                      _ptr(y), _stream(x)), "rv_gemm_rowstat")
    return y


def rowdot(a: torch.Tensor, b: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """fp32 [rows]: scale * sum_c a[row][c] * b[row][c] for two bf16 matrices (unit stride in c)."""
    _need_cuda(a, b)
    if a.shape != b.shape or a.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1 or a.dtype != torch.bfloat16 or b.dtype != a.dtype:
        raise ValueError("rowdot: two bf16 [rows][cols] matrices with unit column stride")
    out = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
    check(_lib.load().rv_rowdot(_ptr(a), _ptr(b), a.shape[0], a.shape[1], a.stride(0), b.stride(0), float(scale), _ptr(out),
                                _stream(a)), "rv_rowdot")
    return out


def softmax_rows(s: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    _need_cuda(s)
    rows, cols = s.shape
    p = torch.empty((rows, cols), dtype=out_dtype, device=s.device)
    check(_lib.load().rv_softmax_rows(_ptr(s), _ptr(p), rows, cols, s.stride(0), p.stride(0), _dt(p), _stream(s)),
          "rv_softmax_rows")
    return p


# ------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------
def nchw_to_nhwc(x: torch.Tensor, c_pad: int, out_dtype: torch.dtype, scale: float = 1.0, shift: float = 0.0):
    _need_cuda(x)
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((n, h, w, c_pad), dtype=out_dtype, device=x.device)
    check(_lib.load().rv_nchw_to_nhwc(_ptr(x), _ptr(y), n, c, h * w, c_pad, _dt(x), _dt(y), scale, shift, _stream(x)),
          "rv_nchw_to_nhwc")
    return y


def nchw_to_nhwc_hpack(x: torch.Tensor, scale: float = 1.0, shift: float = 0.0) -> torch.Tensor:
    """(N,C,H,W), 3C <= 16 -> (N,H,W,16) bf16: per pixel its three horizontal neighbours ([dx][c], zero padded)."""
    _need_cuda(x)
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((n, h, w, 16), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().rv_nchw_to_nhwc_hpack(_ptr(x), _ptr(y), n, c, h, w, _dt(x), scale, shift, _stream(x)), "rv_nchw_to_nhwc_hpack")
    return y


def nhwc_to_nchw(x: torch.Tensor, c: int, out_dtype: torch.dtype):
    _need_cuda(x)
    n, h, w, cs = x.shape
    y = torch.empty((n, c, h, w), dtype=out_dtype, device=x.device)
    check(_lib.load().rv_nhwc_to_nchw(_ptr(x), _ptr(y), n, c, h * w, cs, _dt(x), _dt(y), _stream(x)), "rv_nhwc_to_nchw")
    return y


# ------------------------------------------------------------------------------------------
# posterior, loss, validation metrics
# ------------------------------------------------------------------------------------------
def reparam(moments: torch.Tensor, noise: Optional[torch.Tensor], want_kl: bool = False, z_shift: float = 0.0,
            z_scale: float = 1.0):
    """moments [N, 2Z, H, W]; returns (z or None, kl or None)."""
    _need_cuda(moments, noise)
    moments = moments.contiguous()
    n, c2, h, w = moments.shape
    zc = c2 // 2
    z = None
    if noise is not None:
        if tuple(noise.shape) != (n, zc, h, w):
            raise ValueError(f"noise shape {tuple(noise.shape)} does not match latent shape {(n, zc, h, w)}")
        noise = noise.to(moments.dtype).contiguous()
        z = torch.empty((n, zc, h, w), dtype=moments.dtype, device=moments.device)
    kl = torch.empty((n,), dtype=torch.float32, device=moments.device) if want_kl else None
    check(_lib.load().rv_reparam(_ptr(moments), _ptr(noise), _ptr(z), _ptr(kl), n, zc, h * w, _dt(moments), z_shift,
                                 z_scale, _stream(moments)), "rv_reparam")
    return z, kl


def _pair(a: torch.Tensor, b: torch.Tensor):
    _need_cuda(a, b)
    if a.shape != b.shape or a.dim() != 4 or a.shape[1] != 4:
        raise ValueError(f"expected two (B,4,H,W) tensors, got {tuple(a.shape)} and {tuple(b.shape)}")
    if b.dtype != a.dtype:
        b = b.to(a.dtype)
    return a.contiguous(), b.contiguous()


def recon_loss_per_sample(pred: torch.Tensor, target: torch.Tensor, eb: Sequence[float], eb2: Sequence[float],
                          naive_mse: bool = False) -> torch.Tensor:
    """Per-sample SUM of the AlphaVAE reconstruction loss map, fp32 [B]."""
    pred, target = _pair(pred, target)
    n, _, h, w = pred.shape
    lib = _lib.load()
    blocks = lib.rv_reduce_blocks(h * w)
    partial = torch.empty((n * blocks,), dtype=torch.float64, device=pred.device)
    out = torch.empty((n,), dtype=torch.float32, device=pred.device)
    ebv = (C.c_float * 3)(*[float(v) for v in eb])
    eb2v = (C.c_float * 3)(*[float(v) for v in eb2])
    check(lib.rv_recon_loss(_ptr(pred), _ptr(target), ebv, eb2v, int(naive_mse), _ptr(out), _ptr(partial), n, h * w,
                            _dt(pred), _stream(pred)), "rv_recon_loss")
    return out


def composite_psnr(recon: torch.Tensor, target: torch.Tensor, backgrounds: Sequence[Sequence[float]]) -> torch.Tensor:
    """[B, len(backgrounds)+1] fp32: PSNR (dB) of the composites over each RGB background, then alpha MAE."""
    recon, target = _pair(recon, target)
    n, _, h, w = recon.shape
    nbg = len(backgrounds)
    if nbg > 4:
        raise ValueError("at most 4 backgrounds per call")
    lib = _lib.load()
    blocks = lib.rv_reduce_blocks(h * w)
    partial = torch.empty((n * blocks * 5,), dtype=torch.float64, device=recon.device)
    out = torch.empty((n, nbg + 1), dtype=torch.float32, device=recon.device)
    flat = [float(v) for bg in backgrounds for v in bg]
    bgs = (C.c_float * max(1, len(flat)))(*flat)
    check(lib.rv_composite_psnr(_ptr(recon), _ptr(target), bgs, nbg, _ptr(out), _ptr(partial), n, h * w, _dt(recon),
                                _stream(recon)), "rv_composite_psnr")
    return out


def rgba_loss_terms(recon: torch.Tensor, target: torch.Tensor, eb: Sequence[float], eb2: Sequence[float]) -> torch.Tensor:
    """[B, 6] fp32 per-sample sums of every term ``RgbaVAE.loss`` weighs (rgba_vae.py:283-316): AlphaVAE map on the
    rescaled pair, rgb squared error, white / black composite squared error, alpha squared error, alpha absolute error."""
    recon, target = _pair(recon, target)
    n, _, h, w = recon.shape
    lib = _lib.load()
    blocks = lib.rv_reduce_blocks(h * w)
    partial = torch.empty((n * blocks * 6,), dtype=torch.float64, device=recon.device)
    out = torch.empty((n, 6), dtype=torch.float32, device=recon.device)
    ebv = (C.c_float * 3)(*[float(v) for v in eb])
    eb2v = (C.c_float * 3)(*[float(v) for v in eb2])
    check(lib.rv_rgba_loss_terms(_ptr(recon), _ptr(target), ebv, eb2v, _ptr(out), _ptr(partial), n, h * w, _dt(recon),
                                 _stream(recon)), "rv_rgba_loss_terms")
    return out


def kl_to_reference(moments: torch.Tensor, ref_moments: torch.Tensor, grad_weight: Optional[float] = None):
    """Per-sample KL(posterior || reference posterior) [n] (fp32) and, with ``grad_weight``, grad_weight * d KL / d moments
    (rgba_vae_stage.py:489-508; DiagonalGaussianDistribution.kl(other))."""
    _need_cuda(moments, ref_moments)
    moments = moments.contiguous()
    ref_moments = ref_moments.to(moments.dtype).contiguous()
    if moments.shape != ref_moments.shape:
        raise ValueError(f"posteriors differ in shape: {tuple(moments.shape)} vs {tuple(ref_moments.shape)}")
    n, c2, h, w = moments.shape
    kl = torch.zeros(n, dtype=torch.float32, device=moments.device)
    dm = torch.empty_like(moments) if grad_weight is not None else None
    check(_lib.load().rv_kl_ref(_ptr(moments), _ptr(ref_moments), _ptr(kl), _ptr(dm), n, c2 // 2, h * w, _dt(moments),
                                0.0 if grad_weight is None else float(grad_weight), _stream(moments)), "rv_kl_ref")
    return kl, dm


def blend_tiles(a: torch.Tensor, b: torch.Tensor, extent: int, vertical: bool) -> torch.Tensor:
    """diffusers blend_v / blend_h on contiguous NCHW tiles; b is modified in place and returned."""
    _need_cuda(a, b)
    if not (a.is_contiguous() and b.is_contiguous()) or a.dtype != b.dtype or a.shape[:2] != b.shape[:2]:
        raise ValueError("blend_tiles needs two contiguous NCHW tiles of the same dtype, batch and channels")
    planes = a.shape[0] * a.shape[1]
    check(_lib.load().rv_blend_tiles(_ptr(a), _ptr(b), planes, a.shape[2], a.shape[3], b.shape[2], b.shape[3], int(extent),
                                     int(vertical), _dt(a), _stream(a)), "rv_blend_tiles")
    return b


def attn_scale(c: int) -> float:
    return 1.0 / math.sqrt(c)
