"""Building blocks of the ``rgba_vae`` training step (reference src/training/rgba_vae_stage.py:433-523) on the GPU.

What is here: backward of the AlphaVAE reconstruction loss, of the posterior sample (+KL), of RMS-norm + SiLU, the
data gradient of the stride-1 convolutions (the forward tcgen05 kernels run on flipped/transposed weights), the
flat-buffer AdamW with fused gradient scaling and ``clip_grad_norm_``, and the bucketed data-parallel gradient
all-reduce (NCCL over NVLink on the GPUs; the only collective of the whole path), the weight gradient of the
convolutions and the spatial helpers of the strided / up-sampling layers.  ``trainer.VaeTrainStep`` strings them into
the full step.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import RV_BF16, check
from .ops import _dt, _need_cuda, _ptr, _stream


# ------------------------------------------------------------------------------------------
# backward of the non-conv pieces
# ------------------------------------------------------------------------------------------
def recon_loss_backward(pred: torch.Tensor, target: torch.Tensor, eb: Sequence[float], eb2: Sequence[float],
                        reduce_mean: bool = False, naive_mse: bool = False, grad_output: float = 1.0,
                        clamp: Optional[Sequence[float]] = None) -> torch.Tensor:
    """d reconstruction_loss / d pred (losses.py:67-83 with the reduce rule of :117-123).  ``clamp=(lo, hi)``: pred is
    the decoder's clamped output; the gradient is zero where it sits on a bound (backward of torch.clamp)."""
    pred, target = ops._pair(pred, target)
    n, _, h, w = pred.shape
    ch = 4 if naive_mse else 3
    scale = grad_output / (n * ch * h * w) if reduce_mean else grad_output / n
    out = torch.empty_like(pred)
    ebv = (C.c_float * 3)(*[float(v) for v in eb])
    eb2v = (C.c_float * 3)(*[float(v) for v in eb2])
    lo, hi = (0.0, 0.0) if clamp is None else (float(clamp[0]), float(clamp[1]))
    check(_lib.load().rv_recon_loss_bwd(_ptr(pred), _ptr(target), ebv, eb2v, int(naive_mse), scale, lo, hi, _ptr(out), n, h * w,
                                        _dt(pred), _stream(pred)), "rv_recon_loss_bwd")
    return out


def reparam_backward(moments: torch.Tensor, noise: Optional[torch.Tensor], dz: Optional[torch.Tensor],
                     kl_weight: float = 0.0) -> torch.Tensor:
    """d / d moments of  <dz, posterior.sample(noise)> + kl_weight * sum_b posterior.kl()[b]."""
    _need_cuda(moments, noise, dz)
    moments = moments.contiguous()
    n, c2, h, w = moments.shape
    if (dz is None) != (noise is None):
        raise ValueError("dz and noise must be given together")
    if dz is not None:
        dz, noise = dz.to(moments.dtype).contiguous(), noise.to(moments.dtype).contiguous()
    out = torch.empty_like(moments)
    check(_lib.load().rv_reparam_bwd(_ptr(moments), _ptr(noise), _ptr(dz), _ptr(out), n, c2 // 2, h * w, _dt(moments),
                                     float(kl_weight), _stream(moments)), "rv_reparam_bwd")
    return out


kl_to_reference = ops.kl_to_reference  # lives with the other posterior ops; kept here for the trainer's imports


def rmsnorm_silu_backward(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, silu: bool = True,
                          dgamma_out: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None):
    """Backward of ops.rmsnorm_silu: returns (dx, dgamma) with dgamma shaped like ``gamma`` (fp32).  ``dgamma_out``: a
    contiguous fp32 buffer of C elements (e.g. a view of the flat gradient) the gradient is ACCUMULATED into.  ``add``: the
    gradient of the skip branch that meets the normalised one here; it is added to dx in the same pass."""
    _need_cuda(x, gamma, dy)
    c = x.shape[-1]
    x, dy = x.contiguous(), dy.to(x.dtype).contiguous()
    g_scaled = (gamma.detach().to(torch.float32).reshape(-1) * math.sqrt(c)).contiguous()
    dx = torch.empty_like(x)
    dg = torch.zeros(c, dtype=torch.float32, device=x.device) if dgamma_out is None else dgamma_out
    if add is not None:
        add = add.to(x.dtype).contiguous()
    check(_lib.load().rv_rmsnorm_silu_bwd(_ptr(x), _ptr(g_scaled), _ptr(dy), _ptr(add), _ptr(dx), _ptr(dg), math.sqrt(c), x.numel() // c, c,
                                          _dt(x), int(silu), _stream(x)), "rv_rmsnorm_silu_bwd")
    return dx, dg.reshape(gamma.shape)


def groupnorm_silu_backward(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, dy: torch.Tensor,
                            groups: int = 32, eps: float = 1e-6, silu: bool = True, dgamma_out: Optional[torch.Tensor] = None,
                            dbeta_out: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None):
    """Backward of ops.groupnorm_silu (``stats`` = its ``return_stats`` output): returns (dx, dgamma, dbeta).  ``dgamma_out`` /
    ``dbeta_out``: contiguous fp32 [C] buffers the parameter gradients are ACCUMULATED into (views of the flat gradient);
    ``add``: the skip branch's gradient, added to dx in the same pass."""
    _need_cuda(x, stats, gamma, beta, dy)
    n, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c)
    x, dy = x.contiguous(), dy.to(x.dtype).contiguous()
    g32 = gamma.detach().to(torch.float32).reshape(-1).contiguous()
    b32 = beta.detach().to(torch.float32).reshape(-1).contiguous()
    dx = torch.empty_like(x)
    dg = torch.zeros(c, dtype=torch.float32, device=x.device) if dgamma_out is None else dgamma_out
    db = torch.zeros(c, dtype=torch.float32, device=x.device) if dbeta_out is None else dbeta_out
    if add is not None:
        add = add.to(x.dtype).contiguous()
    scratch = torch.empty(n * c * 2, dtype=torch.float64, device=x.device)
    check(_lib.load().rv_groupnorm_silu_bwd(_ptr(x), _ptr(stats), _ptr(g32), _ptr(b32), _ptr(dy), _ptr(add), _ptr(dx), _ptr(dg), _ptr(db),
                                            _ptr(scratch), n, hw, c, groups, eps, _dt(x), int(silu), _stream(x)), "rv_groupnorm_silu_bwd")
    return dx, dg, db


def conv_dgrad_weights(weight2d: torch.Tensor, cout_pad: int = 0) -> torch.Tensor:
    """[cout][cin][k][k] (any outer strides, e.g. the last temporal slice of a causal 3-D kernel) -> packed bf16 weights
    of the convolution that maps dY to dX for a stride-1 conv: W'[cin][k-1-dy][k-1-dx][cout (zero padded)]."""
    cout, cin, k = weight2d.shape[0], weight2d.shape[1], weight2d.shape[2]
    w = weight2d.detach()
    if w.dtype != torch.bfloat16 or w.stride(-1) != 1 or (k > 1 and w.stride(-2) != k):
        w = w.to(torch.bfloat16).contiguous()
    cp = max(cout, cout_pad)
    out = torch.empty((cin, k * k * cp), dtype=torch.bfloat16, device=w.device)
    check(_lib.load().rv_pack_dgrad_weights(_ptr(w), w.stride(0), w.stride(1), _ptr(out), cout, cin, cp, k, _stream(w)),
          "rv_pack_dgrad_weights")
    return out


def conv_dgrad(dy: torch.Tensor, weight2d: torch.Tensor, pad_lo: Optional[int] = None) -> torch.Tensor:
    """dX (NHWC bf16) of a stride-1 3x3 pad-1 (or 1x1) convolution from dY (NHWC bf16): the forward tensor-core
    kernel on the flipped, transposed weights (no bias).  dY may carry zero-padded channels beyond the conv's cout.
    ``pad_lo=2`` with a zero-inserted dY is the data gradient of the stride-2 (0,1,0,1)-padded down-sampling conv."""
    _need_cuda(dy, weight2d)
    n, h, w, cy = dy.shape
    cin, k = weight2d.shape[1], weight2d.shape[2]
    wp = conv_dgrad_weights(weight2d, cy)
    dx = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device=dy.device)
    desc = ops.make_desc(n, h, w, cy, cin, k, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16, bias_mode=0)
    if pad_lo is not None:
        desc.pad_lo = pad_lo
    ops.conv2d_tc(desc, dy.contiguous(), wp, wp.shape[1], None, None, dx)
    return dx


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, want_bias: bool = True, pad: Optional[int] = None,
               dw_out: Optional[torch.Tensor] = None, dbias_out: Optional[torch.Tensor] = None):
    """(dW [cout][cin][k][k], dbias [cout]) in fp32 of a stride-1 'same' convolution from x and dY (NHWC bf16).
    ``pad=0`` with a zero-inserted dY gives the gradient of the stride-2 (0,1,0,1)-padded conv.
    ``dw_out`` / ``dbias_out``: fp32 views (e.g. of the flat gradient buffer; dw_out [cout'][cin'][k][k] with arbitrary
    outer strides, cout' <= dY channels, cin' <= x channels) the gradients are ACCUMULATED into."""
    _need_cuda(x, dy)
    n, h, w, cin = x.shape
    cout = dy.shape[-1]
    if dw_out is None:
        dw_out = torch.zeros((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
    if dbias_out is None and want_bias:
        dbias_out = torch.zeros(dw_out.shape[0], dtype=torch.float32, device=x.device)
    if dw_out.dtype != torch.float32 or dw_out.stride(-1) != 1 or (ksize > 1 and dw_out.stride(-2) != ksize):
        raise ValueError("dw_out must be fp32 with dense [k][k] inner dimensions")
    check(_lib.load().rv_conv2d_wgrad(_ptr(x.contiguous()), _ptr(dy.contiguous()), _ptr(dw_out), dw_out.stride(0), dw_out.stride(1), 1,
                                      _ptr(dbias_out), n, h, w, cin, cout, dw_out.shape[1], dw_out.shape[0], ksize,
                                      ksize // 2 if pad is None else pad, _stream(x)), "rv_conv2d_wgrad")
    return dw_out, dbias_out


def gemm_tn_accumulate(dst: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> None:
    """dst[m][n] (fp32, contiguous) += a[rows][m]^T . b[rows][n]  (bf16 row-major operands; the reduction runs over the
    rows, which is the pixel axis of the weight-gradient kernel)."""
    _need_cuda(dst, a, b)
    rows, m = a.shape
    nn_ = b.shape[1]
    check(_lib.load().rv_conv2d_wgrad(_ptr(b), _ptr(a), _ptr(dst), nn_, 1, 0, None, 1, 1, rows, nn_, m, nn_, m, 1, 0, _stream(a)),
          "rv_conv2d_wgrad")


def resample2x(x: torch.Tensor, mode: str) -> torch.Tensor:
    """NHWC bf16: 'zero_insert' (x2, zeros between), 'nearest' (x2), 'sum_pool' (2x2 sums)."""
    _need_cuda(x)
    m = {"zero_insert": 0, "nearest": 1, "sum_pool": 2}[mode]
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c) if m == 2 else (n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    check(_lib.load().rv_resample2x(_ptr(x.contiguous()), _ptr(y), n, h, w, c, m, _stream(x)), "rv_resample2x")
    return y


def add_(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b (bf16, same shape) through librgbavae."""
    _need_cuda(a, b)
    y = torch.empty_like(a)
    check(_lib.load().rv_add_bf16(_ptr(a.contiguous()), _ptr(b.contiguous()), _ptr(y), a.numel(), _stream(a)), "rv_add_bf16")
    return y


# ------------------------------------------------------------------------------------------
# optimizer + gradient all-reduce
# ------------------------------------------------------------------------------------------
class FlatAdamW:
    """AdamW (torch.optim.AdamW semantics; rgba_vae_stage.py:321-331 uses betas (0.5, 0.9), lr 1e-5) over ONE flat fp32
    master buffer.  Parameters become views of a flat buffer in the model dtype, gradients are written into the flat
    fp32 ``grad`` buffer; ``step`` = squared-norm kernel + one fused clip/scale/update kernel."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-5, betas=(0.5, 0.9), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: Optional[float] = 1.0):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        _need_cuda(self.params[0])
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in self.sizes:
            self.offsets.append(self.offsets[-1] + s)
        n = self.offsets[-1]
        self.master = torch.empty(n, dtype=torch.float32, device=dev)
        for p, o, s in zip(self.params, self.offsets, self.sizes):
            self.master[o:o + s].copy_(p.detach().reshape(-1))
        self.model_flat = self.master if dt == torch.float32 else self.master.to(dt)
        for p, o, s in zip(self.params, self.offsets, self.sizes):  # parameters now alias the flat buffer
            p.data = self.model_flat[o:o + s].view(p.shape)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(self.master)
        self.v = torch.zeros_like(self.master)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._sq_scratch = torch.zeros(_lib.load().rv_grad_sqnorm_scratch_bytes(), dtype=torch.uint8, device=dev)
        # [t, 1 - beta1^t, 1 - beta2^t] on the device: the step is replayable from a CUDA graph
        self.state = torch.zeros(3, dtype=torch.float32, device=dev)

    @property
    def t(self) -> int:
        return int(self.state[0].item())

    def reset_state(self) -> None:
        """Back to step 0 with zero moments (the master weights are kept)."""
        self.state.zero_()
        self.m.zero_()
        self.v.zero_()

    def grad_view(self, i: int) -> torch.Tensor:
        return self.grad[self.offsets[i]:self.offsets[i + 1]].view(self.params[i].shape)

    def zero_grad(self) -> None:
        self.grad.zero_()

    def step(self, grad_scale: float = 1.0) -> None:
        """``grad_scale`` = 1/world_size after a SUM all-reduce."""
        lib = _lib.load()
        st = _stream(self.master)
        check(lib.rv_adamw_advance(_ptr(self.state), self.betas[0], self.betas[1], st), "rv_adamw_advance")
        sq = None
        mx = 0.0
        if self.max_grad_norm is not None and self.max_grad_norm > 0:
            self.sqnorm.zero_()
            check(lib.rv_grad_sqnorm(_ptr(self.grad), self.grad.numel(), _ptr(self.sqnorm), _ptr(self._sq_scratch), st), "rv_grad_sqnorm")
            sq, mx = _ptr(self.sqnorm), float(self.max_grad_norm)
        pb = None if self.model_flat is self.master else _ptr(self.model_flat)
        check(lib.rv_adamw_step(_ptr(self.master), _ptr(self.grad), _ptr(self.m), _ptr(self.v), pb, self.master.numel(),
                                self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, 0, _ptr(self.state), grad_scale,
                                sq, mx, st), "rv_adamw_step")


class GradientAllReducer:
    """Data-parallel gradient synchronisation: the flat gradient buffer is cut into ``num_buckets`` contiguous buckets
    in REVERSE parameter order (the decoder's gradients are produced first by backward), each bucket is all-reduced
    (SUM) asynchronously as soon as it is marked ready, and ``wait`` joins them before the optimizer step.  On the GPUs
    this is one NCCL all-reduce per bucket over NVLink/NVSwitch; the division by the world size is fused into AdamW."""

    def __init__(self, flat_grad: torch.Tensor, num_buckets: int = 4, group=None, boundary=None):
        """``boundary``: offset(s) in the flat buffer where a group of parameters begins whose gradients are complete EARLIER
        in the backward pass than those before it -- for a VAE the decoder side, or [deep encoder, decoder] (ascending).
        Bucket edges never straddle a boundary: the segments are cut separately (last segment first), every segment gets
        at least one bucket and the rest go in proportion to the sizes, so each segment's buckets can start as soon as its
        part of the backward is done."""
        self.grad, self.group = flat_grad, group
        n = flat_grad.numel()
        num_buckets = max(1, min(num_buckets, n))

        def cut(lo, hi, k):  # k contiguous buckets of [lo, hi), last elements first
            edges = [hi - ((hi - lo) * i) // k for i in range(k + 1)]
            return [(edges[i + 1], edges[i]) for i in range(k)]

        bounds = [] if boundary is None else ([boundary] if isinstance(boundary, int) else list(boundary))
        bounds = sorted(b for b in set(bounds) if 0 < b < n)
        if not bounds or num_buckets < 2:
            self.buckets = cut(0, n, num_buckets)  # last parameters first
        else:
            bounds = bounds[-(num_buckets - 1):]        # at most one segment per bucket
            segs = list(zip([0] + bounds, bounds + [n]))  # ascending; processed last -> first
            k = [1] * len(segs)
            for _ in range(num_buckets - len(segs)):      # hand the spare buckets to the segments with the largest share each
                i = max(range(len(segs)), key=lambda j: (segs[j][1] - segs[j][0]) / k[j])
                k[i] += 1
            self.buckets = [b for (lo, hi), kk in reversed(list(zip(segs, k))) for b in cut(lo, hi, kk)]
        self._work = []
        # measurement (bench.py): with ``timing`` on, ``wait`` brackets every bucket's join with CUDA events on the compute
        # stream; ``exposed_ms()`` then gives, per bucket, how long the compute stream sat blocked on that all-reduce
        self.timing = False
        self._events = []

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def ready(self, bucket: int) -> None:
        lo, hi = self.buckets[bucket]
        if self.world() > 1:
            self._work.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce_all(self) -> None:
        for b in range(len(self.buckets)):
            self.ready(b)

    def wait(self) -> float:
        """Joins outstanding all-reduces; returns the gradient scale (1/world) to hand to ``FlatAdamW.step``."""
        if self.timing and self._work:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(self._work) + 1)]
            ev[0].record()
            for i, w in enumerate(self._work):
                w.wait()
                ev[i + 1].record()
            self._events.append(ev)
        else:
            for w in self._work:
                w.wait()
        self._work = []
        return 1.0 / self.world()

    def bytes_per_step(self) -> int:
        """Payload of one step's all-reduces (fp32 gradients, every bucket once)."""
        return int(self.grad.numel() * self.grad.element_size())

    def exposed_ms(self):
        """Mean exposed time per join (in the order the joins were issued) over the steps recorded with ``timing`` on;
        call after a device synchronize.  Clears the record."""
        if not self._events:
            return []
        n = len(self._events[0]) - 1
        out = [sum(ev[i].elapsed_time(ev[i + 1]) for ev in self._events) / len(self._events) for i in range(n)]
        self._events = []
        return out
