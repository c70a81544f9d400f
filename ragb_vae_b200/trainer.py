"""The ``rgba_vae`` training step (reference src/training/rgba_vae_stage.py:433-523) on the GPU, Qwen-Image arch, bf16.

    inputs -> clamp -> [-1,1] -> detail-augmented triplet -> encode -> posterior.sample -> decode
           -> AlphaVAE reconstruction loss + kl_scale * KL -> backward -> clip_grad_norm_ -> AdamW

There is no autograd graph: the forward pass keeps the activations the hand-written backward needs (a "tape" of plain
tensors per block) and ``backward`` walks the model in reverse, calling librgbavae for every piece:

  * convolutions   data gradient = the forward tcgen05 kernel on flipped / transposed weights; weight gradient =
                   ``rv_conv2d_wgrad`` (MN-major UMMA operands, split-K over pixels, fp32 accumulation).
                   stride-2 down-sampler: dY is zero-inserted onto the input grid (``rv_resample2x``) and both gradients
                   become stride-1 problems; nearest-x2 + conv up-sampler: gradients of the 3x3 conv on the up-sampled
                   grid followed by a 2x2 sum pool.
  * RMS norm+SiLU  ``rv_rmsnorm_silu_bwd``
  * attention      scores are recomputed per block of query rows; dV, dK, dQ are "weight-gradient" GEMMs (the reduction
                   runs over tokens, which is the pixel axis of the wgrad kernel), dP a plain GEMM, and
                   ``rv_softmax_bwd`` writes dS and its transpose.
  * loss / sample  ``rv_recon_loss_bwd`` (with the decoder's [-1,1] clamp mask), ``rv_reparam_bwd`` (+ KL term)
  * optimizer      ``FlatAdamW`` (fused clip + AdamW over one flat buffer); data parallel: bucketed NCCL all-reduce.

What the reference's step also has and this one does not: LPIPS (third-party VGG, out of scope).  The
reference-KL term against a frozen copy (``ref_vae`` + ``ref_kl_scale``; 1e-16 in configs/flux_vae.yaml) is optional: with
it the whole triplet (3B images) runs through the taped encoder and its backward, like the reference's autograd does;
without it the black / white composites are still encoded (``encode_triplet=True``) but carry no gradient.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from . import training as T
from ._lib import RvError
from .autoencoder import RgbaAutoencoder
from .losses import AlphaVaeLoss
from .plumbing import build_detail_augmented_triplet
from .posterior import DiagonalGaussianDistribution


class VaeTrainStep:
    def __init__(self, vae: RgbaAutoencoder, *, lr: float = 1e-5, betas=(0.5, 0.9), eps: float = 1e-8,
                 weight_decay: float = 0.01, max_grad_norm: Optional[float] = 1.0, kl_scale: Optional[float] = 1e-6,
                 loss_module: Optional[AlphaVaeLoss] = None, num_buckets: int = 4, encode_triplet: bool = True, group=None,
                 ref_vae: Optional[RgbaAutoencoder] = None, ref_kl_scale: Optional[float] = None):
        self.flux = vae.arch == "flux"   # GroupNorm blocks, Linear q/k/v/out attention, no quant convs, no decode clamp
        if vae.dtype != torch.bfloat16:
            raise TypeError("the training step runs the model in bfloat16 (fp32 master weights live in the optimizer)")
        self.vae = vae
        self.loss_module = loss_module if loss_module is not None else AlphaVaeLoss()
        self.kl_scale = kl_scale
        self.encode_triplet = encode_triplet
        # reference-KL term (rgba_vae_stage.py:385-406, 489-508): a frozen copy's posteriors of the black / white composites
        self.ref_vae = ref_vae
        self.ref_kl_scale = ref_kl_scale
        # parameters autograd would give a gradient to: everything except the video-only temporal convs
        named = [(n, p) for n, p in vae.named_parameters() if ".time_conv." not in n]
        # flat-buffer order: encoder side first, decoder side last, so that the tail buckets hold decoder gradients only and
        # their all-reduce starts while the encoder's backward still runs (quant_conv belongs to the encoder's backward)
        named = [np for np in named if not self._decoder_side(np[0])] + [np for np in named if self._decoder_side(np[0])]
        boundary = sum(p.numel() for n, p in named if not self._decoder_side(n))
        self.names = [n for n, _ in named]
        self.opt = T.FlatAdamW([p for _, p in named], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                               max_grad_norm=max_grad_norm)
        self._gview: Dict[int, torch.Tensor] = {id(p): self.opt.grad_view(i) for i, (_, p) in enumerate(named)}
        # The encoder's backward runs from its deep, parameter-heavy stages to the shallow ones: the flat buffer's first
        # parameters (conv_in and the first stages, <= 10 % of the encoder side) form a segment of their own, so that the all-reduce
        # of everything deeper starts while the shallow stages' backward still runs and only that small segment is exposed
        # (measured at 8 GPUs before: 3.5 of 3.8 exposed ms in the first encoder-side bucket).
        self._enc_split, shallow_end = self._encoder_split(named, boundary)
        bounds = [boundary] if shallow_end is None else [shallow_end, boundary]
        self._shallow_end = shallow_end
        self.reducer = T.GradientAllReducer(self.opt.grad, num_buckets=num_buckets, group=group, boundary=bounds)
        self._graphs: Dict[tuple, tuple] = {}
        self.launches_per_replay = 0

    def _encoder_split(self, named, boundary: int):
        """-> (number of leading entries of ``_encoder_items()`` that make the SHALLOW phase of the encoder's backward, end
        offset of their parameters in the flat buffer), or (0, None) where the parameters are not laid out that way."""
        off, ends = 0, {}
        for _, p in named:
            off += p.numel()
            ends[id(p)] = off
        vae = self.vae
        mods = [vae.encoder.conv_in] + [m for _, m in self._encoder_items()]
        cum, best = 0, (0, None)
        for i, m in enumerate(mods):
            ps = [p for p in m.parameters() if id(p) in ends]
            cum += sum(p.numel() for p in ps)
            if cum > 0.10 * boundary:
                break
            if ps and max(ends[id(p)] for p in ps) == cum and i >= 1:  # contiguous prefix [0, cum) of the flat buffer
                best = (i, cum)                                        # i entries of _encoder_items() (mods[0] is the stem)
        return best

    # ---- small helpers ---------------------------------------------------------------------
    def _gw(self, conv) -> torch.Tensor:
        """fp32 gradient view of a conv weight as [cout][cin][k][k] (5-D causal kernels: only the last temporal tap is live)."""
        v = self._gview[id(conv.weight)]
        return v[:, :, -1] if v.dim() == 5 else v

    def _conv(self, x, conv, **kw):
        return self.vae._conv(x, conv, **kw)

    def _norm(self, x, norm, silu=True):
        return self.vae._norm(x, norm, silu)

    def _norm_fwd(self, x, norm, silu=True, stats=None):
        """-> (act(norm(x)), what the backward needs besides x: the GroupNorm statistics, or None for the RMS norm).
        ``stats``: (groups, sums) when the conv that produced x left them (``_Stream.gn_stats``)."""
        if self.flux:
            v = self.vae
            pre = stats[1] if stats is not None and stats[0] == norm.num_groups else None
            return ops.groupnorm_silu(x, v._f32(norm.weight, "gn_w"), v._f32(norm.bias, "gn_b"), norm.num_groups, norm.eps, silu,
                                      return_stats=True, stats=pre)
        return self.vae._norm(x, norm, silu), None

    def _norm_bwd(self, norm, x, dy, silu=True, add=None, aux=None):
        """``add``: gradient of the skip branch meeting this one at x (fused into the norm backward's store)."""
        if self.flux:
            if aux is None:  # checkpointed block: the statistics are recomputed with the forward
                aux = self._norm_fwd(x, norm, silu)[1]
            dx, _, _ = T.groupnorm_silu_backward(x, aux, norm.weight, norm.bias, dy, norm.num_groups, norm.eps, silu,
                                                 dgamma_out=self._gview[id(norm.weight)].view(-1),
                                                 dbeta_out=self._gview[id(norm.bias)].view(-1), add=add)
            return dx
        dx, _ = T.rmsnorm_silu_backward(x, norm.gamma, dy, silu, dgamma_out=self._gview[id(norm.gamma)].view(-1), add=add)
        return dx

    def _conv_bwd(self, conv, x, dy, *, mode: str = "same", need_dx: bool = True):
        """x: the conv's NHWC bf16 input (channel-padded for the stems); dy: NHWC bf16 gradient of its output, channel
        count possibly padded above conv.out_channels with zeros.  dW / dbias accumulate straight into the flat gradient
        buffer (in the parameter's own layout); returns dX (or None)."""
        k = conv.k
        gw, gb = self._gw(conv), self._gview[id(conv.bias)]
        if mode == "down":
            dyg = T.resample2x(dy, "zero_insert")
            T.conv_wgrad(x, dyg, k, pad=0, dw_out=gw, dbias_out=gb)
        elif mode == "up":
            dyg = dy
            T.conv_wgrad(T.resample2x(x, "nearest"), dy, k, dw_out=gw, dbias_out=gb)
        else:
            dyg = dy
            T.conv_wgrad(x, dy, k, dw_out=gw, dbias_out=gb)
        if not need_dx:
            return None
        dx = T.conv_dgrad(dyg, conv.weight2d(), pad_lo=2 if mode == "down" else None)
        if mode == "up":
            dx = T.resample2x(dx, "sum_pool")
        return dx

    # ---- residual block --------------------------------------------------------------------
    def _res_fwd(self, x, blk, tape: Optional[list]):
        short = getattr(blk, "conv_shortcut", None)
        a, s1 = self._norm_fwd(x, blk.norm1)
        h = x if short is None else self._conv(x, short)
        # conv1 writes its raw output (the norm backward needs it) AND act(norm2(.)) from the same epilogue where one tile holds
        # the pixel's whole channel vector (RMS norm, Cout <= 256); else the norm kernel runs
        st = self.vae._conv_fused(a, blk.conv1, next_norm=(blk.norm2, True), want_raw=True)
        t, s2 = st.raw, None
        if st.act is not None:
            b = st.act
        else:
            b, s2 = self._norm_fwd(t, blk.norm2, stats=st.gn_stats)
        y = self._conv(b, blk.conv2, residual=h)
        if tape is not None:
            # gradient checkpointing (diffusers: per block): keep only the block input, recompute the rest in the backward
            tape.append(("res", blk, (x, None, None, None, None, None) if self.vae.gradient_checkpointing else (x, a, t, b, s1, s2)))
        return y

    def _res_bwd(self, blk, saved, dy):
        x, a, t, b, s1, s2 = saved
        short = getattr(blk, "conv_shortcut", None)
        if a is None:  # checkpointed block: same kernels, same bits as the forward
            a, s1 = self._norm_fwd(x, blk.norm1)
            st = self.vae._conv_fused(a, blk.conv1, next_norm=(blk.norm2, True), want_raw=True)
            t = st.raw
            if st.act is not None:
                b = st.act
            else:
                b, s2 = self._norm_fwd(t, blk.norm2, stats=st.gn_stats)
        db = self._conv_bwd(blk.conv2, b, dy)
        dt = self._norm_bwd(blk.norm2, t, db, aux=s2)
        da = self._conv_bwd(blk.conv1, a, dt)
        skip = dy if short is None else self._conv_bwd(short, x, dy)
        return self._norm_bwd(blk.norm1, x, da, add=skip, aux=s1)

    # ---- attention -------------------------------------------------------------------------
    def _gemm(self, *a, **k):
        return self.vae._gemm(*a, **k)

    def _attn_weights(self, attn):
        """(Wqkv bf16 [3C, C], bqkv fp32 [3C], Wo bf16 [C, C], bo fp32 [C]) -- Qwen: one 1x1 conv C -> 3C and a 1x1 proj;
        Flux: three Linear(C, C) stacked, and to_out[0]."""
        if self.flux:
            lins = (attn.to_q, attn.to_k, attn.to_v)
            wqkv = torch.cat([l.weight.detach() for l in lins], 0).to(torch.bfloat16).contiguous()
            bqkv = torch.cat([l.bias.detach() for l in lins], 0).to(torch.float32).contiguous()
            out = attn.to_out[0]
            return wqkv, bqkv, out.weight.detach().to(torch.bfloat16).contiguous(), out.bias.detach().to(torch.float32).contiguous()
        c = attn.proj.in_channels
        wqkv = attn.to_qkv.weight.detach().reshape(3 * c, c).to(torch.bfloat16).contiguous()
        bqkv = attn.to_qkv.bias.detach().to(torch.float32).contiguous()
        wo = attn.proj.weight.detach().reshape(c, c).to(torch.bfloat16).contiguous()
        bo = attn.proj.bias.detach().to(torch.float32).contiguous()
        return wqkv, bqkv, wo, bo

    def _attn_norm(self, attn):
        return attn.group_norm if self.flux else attn.norm

    def _attn_fwd(self, x, attn, tape: Optional[list]):
        n, h, w, c = x.shape
        t = h * w
        dev = x.device
        xn, sn = self._norm_fwd(x, self._attn_norm(attn), silu=False)
        wqkv, bqkv, wo, bo = self._attn_weights(attn)
        xn2 = xn.view(n * t, c)
        q = torch.empty((n * t, c), dtype=torch.bfloat16, device=dev)
        k = torch.empty_like(q)
        v = torch.empty_like(q)
        for i, dst in enumerate((q, k, v)):
            self._gemm(xn2, wqkv[i * c:(i + 1) * c], rows=n * t, k=c, cols=c, x_ld=c, w_ld=c, y=dst, y_ld=c,
                       bias=bqkv[i * c:(i + 1) * c].contiguous(), bias_mode=1)
        scale = ops.attn_scale(c)
        if self.vae.fused_attention and c in ops.FUSED_ATTENTION_DIMS and t % 128 == 0:
            # flash kernel: scores / probabilities never reach HBM in the forward (the backward recomputes them per block)
            vt_all = torch.empty((c, n * t), dtype=torch.bfloat16, device=dev)  # V^T of all images from one GEMM
            self._gemm(wqkv[2 * c:], xn2, rows=c, k=c, cols=n * t, x_ld=c, w_ld=c, y=vt_all, y_ld=n * t,
                       bias=bqkv[2 * c:].contiguous(), bias_mode=2)
            o, lse = ops.attention(q, k, vt_all, n, t, return_lse=True)
        else:
            lse = None
            o = torch.empty_like(q)
            vt = torch.empty((c, t), dtype=torch.bfloat16, device=dev)
            q_chunk = self._q_chunk(t)
            s = torch.empty((q_chunk, t), dtype=torch.float32, device=dev)
            for i in range(n):
                sl = slice(i * t, (i + 1) * t)
                # V^T[c][token] (the PV GEMM wants the reduction axis contiguous)
                self._gemm(wqkv[2 * c:], xn2[sl], rows=c, k=c, cols=t, x_ld=c, w_ld=c, y=vt, y_ld=t, bias=bqkv[2 * c:].contiguous(),
                           bias_mode=2)
                for r0 in range(0, t, q_chunk):
                    rows = min(q_chunk, t - r0)
                    self._gemm(q[i * t + r0:i * t + r0 + rows], k[sl], rows=rows, k=c, cols=t, x_ld=c, w_ld=c, y=s[:rows], y_ld=t,
                               alpha=scale)
                    p = ops.softmax_rows(s[:rows], torch.bfloat16)
                    self._gemm(p, vt, rows=rows, k=t, cols=c, x_ld=t, w_ld=t, y=o[i * t + r0:i * t + r0 + rows], y_ld=c)
        out = torch.empty_like(x)
        self._gemm(o, wo, rows=n * t, k=c, cols=c, x_ld=c, w_ld=c, y=out.view(n * t, c), y_ld=c, bias=bo, bias_mode=1,
                   residual=x.view(n * t, c))
        if tape is not None:
            tape.append(("attn", attn, (x,) if self.vae.gradient_checkpointing else (x, xn, q, k, v, o, sn, lse)))
        return out

    @staticmethod
    def _q_chunk(t: int) -> int:
        return max(64, min(t, ((1 << 27) // t) // 64 * 64))

    def _attn_bwd(self, attn, saved, dout):
        if len(saved) == 1:  # checkpointed block: recompute the forward with a scratch tape to get its intermediates back
            scratch: list = []
            ckpt, self.vae.gradient_checkpointing = self.vae.gradient_checkpointing, False
            try:
                self._attn_fwd(saved[0], attn, scratch)
            finally:
                self.vae.gradient_checkpointing = ckpt
            saved = scratch[0][2]
        x, xn, q, k, v, o, sn, lse = saved
        n, h, w, c = x.shape
        t = h * w
        dev = x.device
        scale = ops.attn_scale(c)
        # [rows][ch] as an NHWC image for the 1x1-conv forms of the projections' gradients: any pixel arrangement is equivalent,
        # but the weight-gradient kernel splits its reduction over IMAGE ROWS -- as one 65 536-pixel row it ran on 9-27 CTAs
        # (0.7 ms per projection at 54-165 TFLOP/s), as 128-pixel rows it fills the GPU
        def as_img(a, ch):
            rows = a.shape[0]
            return a.view(1, rows // 128, 128, ch) if rows % 128 == 0 else a.view(1, 1, rows, ch)
        dout2 = dout.view(n * t, c)
        # proj: out = o Wo^T + bo + x
        proj = attn.to_out[0] if self.flux else attn.proj
        T.conv_wgrad(as_img(o, c), as_img(dout2, c), 1, dw_out=self._gview[id(proj.weight)].view(c, c, 1, 1),
                     dbias_out=self._gview[id(proj.bias)])
        d_o = T.conv_dgrad(as_img(dout2, c), proj.weight.detach().reshape(c, c, 1, 1)).view(n * t, c)
        dqkv = torch.empty((n * t, 3 * c), dtype=torch.bfloat16, device=dev)
        # With the forward's log-sum-exp the probabilities come straight out of the QK^T GEMM's epilogue (exp2(s - lse)) and dS
        # out of the dO V^T GEMM's (P * (dP - delta), delta = rowsum(dO * O)): no fp32 score matrix, no softmax pass.
        fused = lse is not None
        q_chunk = max(128, min(t, ((1 << 28) // t) // 128 * 128)) if fused else self._q_chunk(t)
        ds = torch.empty((q_chunk, t), dtype=torch.bfloat16, device=dev)
        if fused:
            pbuf = torch.empty((q_chunk, t), dtype=torch.bfloat16, device=dev)
            delta = ops.rowdot(d_o, o, scale)
            scale_log2 = scale * 1.4426950408889634
        else:
            s = torch.empty((q_chunk, t), dtype=torch.float32, device=dev)
            dp = torch.empty((q_chunk, t), dtype=torch.float32, device=dev)
        wqkv, bqkv, _, _ = self._attn_weights(attn)
        xn2 = xn.view(n * t, c)
        # K^T[c][image * token] for all images from one GEMM (dQ = dS K wants the key index contiguous in its second operand)
        kt_all = torch.empty((c, n * t), dtype=torch.bfloat16, device=dev)
        self._gemm(wqkv[c:2 * c], xn2, rows=c, k=c, cols=n * t, x_ld=c, w_ld=c, y=kt_all, y_ld=n * t, bias=bqkv[c:2 * c].contiguous(),
                   bias_mode=2)
        from . import _lib
        for i in range(n):
            sl = slice(i * t, (i + 1) * t)
            dv = torch.zeros((t, c), dtype=torch.float32, device=dev)
            dk = torch.zeros((t, c), dtype=torch.float32, device=dev)
            kt = kt_all[:, i * t:(i + 1) * t]
            for r0 in range(0, t, q_chunk):
                rows = min(q_chunk, t - r0)
                rs = slice(i * t + r0, i * t + r0 + rows)
                if fused:
                    p = ops.gemm_rowstat(q[rs], k[sl], lse[rs], 1, scale_log2, out=pbuf[:rows])
                    ops.gemm_rowstat(d_o[rs], v[sl], delta[rs], 2, scale, mul_in=p, out=ds[:rows])
                else:
                    self._gemm(q[rs], k[sl], rows=rows, k=c, cols=t, x_ld=c, w_ld=c, y=s[:rows], y_ld=t, alpha=scale)
                    p = ops.softmax_rows(s[:rows], torch.bfloat16)
                    # dP = dO V^T
                    self._gemm(d_o[rs], v[sl], rows=rows, k=c, cols=t, x_ld=c, w_ld=c, y=dp[:rows], y_ld=t)
                    T.check(_lib.load().rv_softmax_bwd(ops._ptr(p), ops._ptr(dp), ops._ptr(ds), None, rows, t, t, r0, scale,
                                                       ops._stream(p)), "rv_softmax_bwd")
                # dV += P^T dO ; dK += dS^T Q   (reduction over this block's query rows: wgrad-form GEMMs)
                T.gemm_tn_accumulate(dv, p, d_o[rs])
                T.gemm_tn_accumulate(dk, ds[:rows], q[rs])
                # dQ = dS K, straight into its third of dqkv
                self._gemm(ds[:rows], kt, rows=rows, k=t, cols=c, x_ld=t, w_ld=n * t, y=dqkv[rs, :c], y_ld=3 * c)
            dqkv[sl, c:2 * c] = dk
            dqkv[sl, 2 * c:] = dv
        if self.flux:  # three Linear(C, C): one stacked [3C, C] weight gradient, then a slice into each parameter's view
            dw = torch.zeros((3 * c, c, 1, 1), dtype=torch.float32, device=dev)
            db = torch.zeros(3 * c, dtype=torch.float32, device=dev)
            T.conv_wgrad(as_img(xn.view(n * t, c), c), as_img(dqkv, 3 * c), 1, dw_out=dw, dbias_out=db)
            for i, lin in enumerate((attn.to_q, attn.to_k, attn.to_v)):
                self._gview[id(lin.weight)].add_(dw[i * c:(i + 1) * c].view(c, c))
                self._gview[id(lin.bias)].add_(db[i * c:(i + 1) * c])
            dxn = T.conv_dgrad(as_img(dqkv, 3 * c), wqkv.view(3 * c, c, 1, 1)).view(x.shape)
        else:
            T.conv_wgrad(as_img(xn.view(n * t, c), c), as_img(dqkv, 3 * c), 1, dw_out=self._gview[id(attn.to_qkv.weight)],
                         dbias_out=self._gview[id(attn.to_qkv.bias)])
            dxn = T.conv_dgrad(as_img(dqkv, 3 * c), attn.to_qkv.weight.detach().reshape(3 * c, c, 1, 1)).view(x.shape)
        return self._norm_bwd(self._attn_norm(attn), x, dxn, silu=False, add=dout, aux=sn)

    # ---- encoder / decoder -----------------------------------------------------------------
    def _run_fwd(self, x, items, tape):
        for kind, m in items:
            if kind == "res":
                x = self._res_fwd(x, m, tape)
            elif kind == "attn":
                x = self._attn_fwd(x, m, tape)
            elif kind == "down":
                y = self._conv(x, m)
                if tape is not None:
                    tape.append(("down", m, x))
                x = y
            elif kind == "up":
                y = self._conv(x, m, upsample=True)
                if tape is not None:
                    tape.append(("up", m, x))
                x = y
            elif kind == "conv":
                y = self._conv(x, m)
                if tape is not None:
                    tape.append(("conv", m, x))
                x = y
            elif kind == "norm":
                y, aux = self._norm_fwd(x, m)
                if tape is not None:
                    tape.append(("norm", m, (x, aux)))
                x = y
            else:
                raise AssertionError(kind)
        return x

    def _run_bwd(self, tape: list, dy, stop: int = 0):
        """Walks the tape backwards down to (not including) its first ``stop`` entries."""
        while len(tape) > stop:
            kind, m, saved = tape.pop()
            if kind == "res":
                dy = self._res_bwd(m, saved, dy)
            elif kind == "attn":
                dy = self._attn_bwd(m, saved, dy)
            elif kind in ("down", "up"):
                dy = self._conv_bwd(m, saved, dy, mode=kind)
            elif kind == "conv":
                dy = self._conv_bwd(m, saved, dy)
            elif kind == "norm":
                dy = self._norm_bwd(m, saved[0], dy, aux=saved[1])
            elif kind == "stem":  # first conv of a network: input has no gradient (or it is returned as is)
                conv, need_dx = m
                dy = self._conv_bwd(conv, saved, dy, need_dx=need_dx)
            else:
                raise AssertionError(kind)
        return dy

    def _encoder_items(self):
        enc = self.vae.encoder
        items = []
        if self.flux:
            for blk in enc.down_blocks:
                items += [("res", r) for r in blk.resnets]
                if getattr(blk, "downsamplers", None) is not None:
                    items.append(("down", blk.downsamplers[0].conv))
            mid = enc.mid_block
            return items + [("res", mid.resnets[0]), ("attn", mid.attentions[0]), ("res", mid.resnets[1]), ("norm", enc.conv_norm_out)]
        for blk in enc.down_blocks:
            items.append(("res", blk) if blk._kind == "res" else ("down", blk.resample[1]))
        mid = enc.mid_block
        items += [("res", mid.resnets[0]), ("attn", mid.attentions[0]), ("res", mid.resnets[1]), ("norm", enc.norm_out),
                  ("conv", enc.conv_out)]
        return items

    def _decoder_items(self):
        dec = self.vae.decoder
        mid = dec.mid_block
        items = [] if self.flux else [("conv", dec.conv_in)]
        items += [("res", mid.resnets[0]), ("attn", mid.attentions[0]), ("res", mid.resnets[1])]
        for blk in dec.up_blocks:
            items += [("res", r) for r in blk.resnets]
            if getattr(blk, "upsamplers", None) is not None:
                up = blk.upsamplers[0]
                items.append(("up", up.conv if self.flux else up.resample[1]))
        items.append(("norm", dec.conv_norm_out if self.flux else dec.norm_out))
        return items

    def encode_moments(self, x_vae: torch.Tensor, tape: Optional[list]) -> torch.Tensor:
        """(B,4,H,W) in [-1,1] -> fp32 moments (B,32,H/8,W/8), recording the tape."""
        vae = self.vae
        vae._check_image(x_vae, 4, "training input", 8)
        xp = ops.nchw_to_nhwc(x_vae.contiguous(), 16, torch.bfloat16)
        y = self._conv(xp, vae.encoder.conv_in)
        if tape is not None:
            tape.append(("stem", (vae.encoder.conv_in, False), xp))
        h = self._run_fwd(y, self._encoder_items(), tape)
        last = vae.encoder.conv_out if self.flux else vae.quant_conv   # Flux: conv_out yields the moments (no quant conv)
        if tape is not None:
            tape.append(("conv", last, h))
        return self._conv(h, last, y_nchw=True, y_dtype=torch.float32)

    def decode(self, z: torch.Tensor, tape: Optional[list]) -> torch.Tensor:
        """fp32 latents (B,16,h,w) -> fp32 image (B,4,8h,8w); Qwen: clamped to [-1,1] (AutoencoderKLQwenImage._decode)."""
        vae = self.vae
        zp = ops.nchw_to_nhwc(z.contiguous(), 16, torch.bfloat16)
        first = vae.decoder.conv_in if self.flux else vae.post_quant_conv
        y = self._conv(zp, first)
        if tape is not None:
            tape.append(("stem", (first, True), zp))
        h = self._run_fwd(y, self._decoder_items(), tape)
        if tape is not None:
            tape.append(("conv", vae.decoder.conv_out, h))
        return self._conv(h, vae.decoder.conv_out, y_nchw=True, y_dtype=torch.float32, clamp=self._decode_clamp())

    def _decode_clamp(self):
        return None if self.flux else (-1.0, 1.0)

    # ---- the step --------------------------------------------------------------------------
    def _forward_and_decoder_backward(self, inputs: torch.Tensor, noise: torch.Tensor):
        """Phase 1: everything up to (and including) the decoder's backward.  Returns (loss terms, context of phase 2)."""
        vae = self.vae
        # (the taped convs go through _conv / _conv_fused(want_raw=True), which never drop the raw output)
        B = inputs.shape[0]
        target_vae = torch.clamp(inputs.to(torch.float32), 0.0, 1.0) * 2.0 - 1.0
        enc_tape: list = []
        use_ref = self.ref_vae is not None and self.ref_kl_scale is not None and self.ref_kl_scale > 0.0
        ref_ctx = None
        if use_ref:
            # the black / white posteriors carry a gradient now: the whole triplet goes through the taped encoder
            composed = build_detail_augmented_triplet(target_vae)
            moments_all = self.encode_moments(composed, enc_tape)
            moments = moments_all[:B]
            ref_moments = self.ref_vae._encode_moments(composed)  # frozen copy, inference path, no tape
            ref_ctx = (moments_all, ref_moments)
        else:
            moments = self.encode_moments(target_vae, enc_tape)
            if self.encode_triplet:  # black / white composites: encoded like the reference does, no gradient path
                composed = build_detail_augmented_triplet(target_vae)
                vae._encode_moments(composed[B:])
        post = DiagonalGaussianDistribution(moments)
        z = post.sample(noise=noise)
        dec_tape: list = []
        pred = self.decode(z, dec_tape)
        lm = self.loss_module
        recon = lm.reconstruction_loss(pred, target_vae)
        metrics = {"train/recon": recon}
        total = recon
        kl_w = 0.0
        if self.kl_scale is not None and self.kl_scale > 0.0:
            kl = lm.kl_loss(post)
            metrics["train/kl"] = kl
            total = total + self.kl_scale * kl
            kl_w = self.kl_scale / B  # posterior.kl() is already the per-sample sum; both reduce rules average it over B
        dm_ref = None
        if use_ref:
            # 0.5 * (kl_loss(black, ref_black) + kl_loss(white, ref_white)); kl_loss = per-sample sums averaged over B
            moments_all, ref_moments = ref_ctx
            w = self.ref_kl_scale * 0.5 / B
            kl_bw, dm_ref = T.kl_to_reference(moments_all[B:], ref_moments[B:], grad_weight=w)
            ref_kl = 0.5 * (kl_bw[:B].mean() + kl_bw[B:].mean())
            metrics["train/ref_kl"] = ref_kl
            total = total + self.ref_kl_scale * ref_kl
        metrics["train/loss"] = total
        # ---- backward ----
        self.opt.zero_grad()
        dpred = T.recon_loss_backward(pred, target_vae, lm._eb, lm._eb2, lm.reduce_mean, lm.use_naive_mse, clamp=self._decode_clamp())
        dy = ops.nchw_to_nhwc(dpred, 16, torch.bfloat16)
        dzp = self._run_bwd(dec_tape, dy)  # NHWC [B,h,w,16]
        return metrics, (enc_tape, moments, noise, dzp, kl_w, dm_ref)

    def _encoder_backward(self, ctx):
        """Phase 2: posterior sample / KL backward and the encoder's backward down to its shallow stages (the first
        ``_enc_split`` blocks + the stem, whose tape entries stay for phase 3).  Returns the gradient phase 3 continues from."""
        enc_tape, moments, noise, dzp, kl_w, dm_ref = ctx
        dz = ops.nhwc_to_nchw(dzp, 16, torch.float32)
        dmom = T.reparam_backward(moments.contiguous(), noise, dz, kl_weight=kl_w)
        if dm_ref is not None:  # the triplet's black / white thirds get the reference-KL gradient
            dmom = torch.cat([dmom, dm_ref], dim=0)
        dm = ops.nchw_to_nhwc(dmom, dmom.shape[1], torch.bfloat16)
        # tape = [stem, one entry per _encoder_items() element ..., last conv]
        return self._run_bwd(enc_tape, dm, stop=1 + self._enc_split if self._enc_split else 0)

    def _encoder_backward_shallow(self, ctx, dy) -> None:
        """Phase 3: the rest of the encoder's backward (nothing to do where the encoder is not split)."""
        if self._enc_split:
            self._run_bwd(ctx[0], dy)

    def forward_backward(self, inputs: torch.Tensor, noise: Optional[torch.Tensor] = None, generator=None) -> Dict[str, torch.Tensor]:
        """inputs: (B,4,H,W) in [0,1].  Fills the optimizer's flat gradient buffer (and starts the bucketed all-reduce of
        the decoder's gradients while the encoder's backward still runs); returns the step's loss terms."""
        if not inputs.is_cuda:
            raise RvError("VaeTrainStep runs on CUDA (sm_100a) only; there is no CPU path")
        self._check_untiled(inputs)
        if noise is None:
            b, _, h, w = inputs.shape
            noise = torch.randn((b, int(self.vae.config.latent_channels), h // 8, w // 8), generator=generator, device=inputs.device,
                                dtype=torch.float32)
        metrics, ctx = self._forward_and_decoder_backward(inputs, noise)
        self._mark_ready(0)
        dy = self._encoder_backward(ctx)
        self._mark_ready(1)
        self._encoder_backward_shallow(ctx, dy)
        self._mark_ready(2)
        return metrics

    def _check_untiled(self, inputs: torch.Tensor) -> None:
        """The taped step differentiates the untiled encode / decode.  With ``enable_tiling()`` the reference would tile
        (and seam-blend) any input larger than the tile -- a different function: refuse instead of silently diverging."""
        if self.vae.use_tiling:
            tile = self.vae._tiling()[0]
            if inputs.shape[-1] > tile or inputs.shape[-2] > tile:
                raise NotImplementedError(f"VaeTrainStep does not differentiate the tiled path: input {tuple(inputs.shape[-2:])} "
                                          f"exceeds the {tile}-pixel tile; call vae.disable_tiling() (slicing is a no-op and is fine)")

    def _mark_ready(self, phase: int) -> None:
        """Start the all-reduce of every bucket whose parameters all have their gradients.  Phase 0: the decoder's backward is
        done (its parameters are the tail of the flat buffer, so its buckets go first while the encoder's backward runs);
        1: the encoder's deep stages are done (everything from ``_shallow_end`` on); 2: everything."""
        if self.reducer.world() == 1:
            return
        if phase == 0:
            self._started = set()
        for b, (lo, hi) in enumerate(self.reducer.buckets):
            if b in self._started:
                continue
            ok = phase == 2 or (phase == 0 and self._bucket_is_decoder(lo, hi)) or \
                (phase == 1 and self._shallow_end is not None and lo >= self._shallow_end)
            if ok:
                self.reducer.ready(b)
                self._started.add(b)

    @staticmethod
    def _decoder_side(name: str) -> bool:
        return name.startswith(("decoder.", "post_quant_conv."))

    def _bucket_is_decoder(self, lo: int, hi: int) -> bool:
        for i, n in enumerate(self.names):
            if self.opt.offsets[i + 1] > lo and self.opt.offsets[i] < hi and not self._decoder_side(n):
                return False
        return True

    def step(self, inputs: torch.Tensor, noise: Optional[torch.Tensor] = None, generator=None) -> Dict[str, torch.Tensor]:
        """forward + backward + gradient all-reduce + clip + AdamW; returns the loss terms (device scalars)."""
        metrics = self.forward_backward(inputs, noise, generator)
        scale = self.reducer.wait()
        self.opt.step(grad_scale=scale)
        self.vae.mark_weights_changed()  # packed weights (and captured inference graphs) are stale after the in-place update
        return metrics

    # ---- CUDA-graph replay ------------------------------------------------------------------
    def step_graphed(self, inputs: torch.Tensor, noise: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``step`` as four CUDA graphs captured once per input shape and sharing one memory pool: (1) forward + decoder
        backward, (2) the encoder's backward down to its shallow stages, (2b) those, (3) clip + AdamW -- ~1100 launches with no
        host work in between.  The bucketed NCCL all-reduces are issued eagerly BETWEEN the replays (decoder buckets after graph
        1, the deep encoder's after graph 2, so they overlap the remaining backward on NCCL's stream).  ``noise`` must be supplied (the posterior's eps); the returned loss terms are views of graph 1's
        static outputs, valid until the next call."""
        self._check_untiled(inputs)
        key = (tuple(inputs.shape), inputs.dtype, tuple(noise.shape), noise.dtype, bool(self.vae.gradient_checkpointing))
        g = self._graphs.get(key)
        if g is None:
            sx, sn = inputs.clone(), noise.clone()
            stream = torch.cuda.Stream()
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                # warm-up outside the capture (TMA attributes, allocator pools, NCCL communicators) WITHOUT the update,
                # so that the first replay below is the first optimizer step
                self.forward_backward(sx, sn)
                self.reducer.wait()
            torch.cuda.current_stream().wait_stream(stream)
            torch.cuda.synchronize()
            self.vae._pack_cache.clear()  # the weight-packing kernels must be part of the graphs (weights change every replay)
            l0 = ops.launch_count()
            g1, g2, g2b, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                metrics, ctx = self._forward_and_decoder_backward(sx, sn)
            with torch.cuda.graph(g2, pool=g1.pool()):
                dy = self._encoder_backward(ctx)
            if self._enc_split:
                with torch.cuda.graph(g2b, pool=g1.pool()):
                    self._encoder_backward_shallow(ctx, dy)
            else:
                g2b = None
            with torch.cuda.graph(g3, pool=g1.pool()):
                self.opt.step(grad_scale=1.0 / self.reducer.world())
            self.vae._pack_cache.clear()
            self.launches_per_replay = ops.launch_count() - l0
            g = self._graphs[key] = (g1, g2, g2b, g3, sx, sn, metrics, ctx)
        g1, g2, g2b, g3, sx, sn, metrics, _ = g
        sx.copy_(inputs, non_blocking=True)
        sn.copy_(noise, non_blocking=True)
        g1.replay()
        self._mark_ready(0)
        g2.replay()
        self._mark_ready(1)
        if g2b is not None:
            g2b.replay()
        self._mark_ready(2)
        self.reducer.wait()
        g3.replay()
        # rv_adamw_step rewrites the parameters through raw pointers (no _version bump): anything an eager call between
        # two replays packed (validation, vae.encode ...) is stale now
        self.vae.mark_weights_changed()
        return metrics

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {n: self.opt.grad_view(i) for i, n in enumerate(self.names)}
