#!/usr/bin/env python
"""Per-call CUDA-event timing of one eager rgba_vae training step, grouped by (op, shape) -- development aid.
    python scripts/train_layer_profile.py [size] [batch] [arch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R
from ragb_vae_b200 import ops
from ragb_vae_b200 import training as T
from ragb_vae_b200 import trainer as TR

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
arch = sys.argv[3] if len(sys.argv) > 3 else "qwen"
recs = []


def wrap(mod, name, describe):
    fn = getattr(mod, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        recs.append((name, describe(*a, **k), e0, e1))
        return out

    setattr(mod, name, inner)


def d_conv(desc, x, w, *rest):
    fl = 2.0 * desc.n * desc.oh * desc.ow * desc.cout * desc.cin * desc.ksize ** 2
    return (f"n{desc.n} {desc.h}x{desc.w} {desc.cin}->{desc.cout} k{desc.ksize} s{desc.stride} up{desc.upsample} pad{desc.pad_lo}", fl, 0.0)


def d_wgrad(x, dy, k, **kw):
    n, h, w, ci = x.shape
    co = dy.shape[-1]
    return (f"x{tuple(x.shape)} dy{tuple(dy.shape)} k{k}", 2.0 * n * h * w * ci * co * k * k, 0.0)


el = lambda t: t.numel() * t.element_size()
wrap(ops, "conv2d_tc", d_conv)
wrap(ops, "conv2d_tc_norm", d_conv)
wrap(ops, "conv_out", d_conv)
wrap(T, "conv_wgrad", d_wgrad)
wrap(ops, "rmsnorm_silu", lambda x, *a, **k: (str(tuple(x.shape)), 0.0, 2.0 * el(x)))
wrap(T, "rmsnorm_silu_backward", lambda x, g, dy, *a, **k: (str(tuple(x.shape)), 0.0, (4.0 if k.get("add") is not None else 3.0) * el(x)))
wrap(ops, "groupnorm_silu", lambda x, *a, **k: (str(tuple(x.shape)), 0.0, 3.0 * el(x)))
wrap(T, "groupnorm_silu_backward", lambda x, *a, **k: (str(tuple(x.shape)), 0.0, 5.0 * el(x)))
wrap(T, "resample2x", lambda x, mode: (f"{tuple(x.shape)} {mode}", 0.0, el(x) * (5.0 if mode != "sum_pool" else 1.25)))
wrap(ops, "attention", lambda q, k, vt, n, t, **kw: (f"n{n} t{t} d{vt.shape[1]}", 4.0 * n * t * t * vt.shape[1], 0.0))
wrap(ops, "gemm_rowstat", lambda x, w, st, mode, alpha, **kw: (f"{tuple(x.shape)} x {tuple(w.shape)}^T mode {mode}",
                                                              2.0 * x.shape[0] * x.shape[1] * w.shape[0], 0.0))
wrap(ops, "rowdot", lambda a, b, s=1.0: (str(tuple(a.shape)), 0.0, 2.0 * el(a)))
wrap(ops, "softmax_rows", lambda s, dt: (str(tuple(s.shape)), 0.0, s.numel() * 6.0))
wrap(T, "gemm_tn_accumulate", lambda out, a, b: (f"{tuple(a.shape)}^T {tuple(b.shape)}", 2.0 * a.shape[0] * a.shape[1] * b.shape[1], 0.0))
wrap(ops, "nchw_to_nhwc", lambda x, cp, dt, *a: (str(tuple(x.shape)), 0.0, el(x) * 3))
wrap(ops, "nhwc_to_nchw", lambda x, c, dt: (str(tuple(x.shape)), 0.0, el(x) * 2))
TR.T = T

torch.manual_seed(0)
vae = R.RgbaAutoencoder(arch).to("cuda", torch.bfloat16)
step = TR.VaeTrainStep(vae, loss_module=R.AlphaVaeLoss(reduce_mean=True))
if os.environ.get("QCHUNK"):
    step._q_chunk = lambda t, c=int(os.environ["QCHUNK"]): min(t, c)
x = torch.rand(B, 4, S, S, device="cuda")
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda")
for _ in range(2):
    recs.clear()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    step.step(x, noise)
    t1.record()
torch.cuda.synchronize()
agg = {}
tot = 0.0
for name, (shape, fl, by), e0, e1 in recs:
    ms = e0.elapsed_time(e1)
    tot += ms
    a = agg.setdefault((name, shape), [0.0, 0, 0.0, 0.0])
    a[0] += ms; a[1] += 1; a[2] += fl; a[3] += by
print(f"{arch} B={B} {S}x{S}: eager step {t0.elapsed_time(t1):.1f} ms, bracketed calls {tot:.1f} ms")
for (name, shape), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    rate = f"{a[2]/a[0]/1e9:8.1f} TFLOP/s" if a[2] else f"{a[3]/a[0]/1e6:8.1f} GB/s"
    print(f"{name:24s} {shape:62s} x{a[1]:<3d} {a[0]:8.3f} ms  {rate}")
