#!/usr/bin/env python
"""Per-launch CUDA-event timing of one RgbaVAE step (development aid): wraps ops.* to bracket each
librgbavae call with events on the current stream and prints shape, ms, TFLOP/s or GB/s.
    python scripts/layer_profile.py [arch] [batch] [size]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ragb_vae_b200 as R
from ragb_vae_b200 import ops

arch = sys.argv[1] if len(sys.argv) > 1 else "qwen"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
recs = []

def wrap(name, fn, describe):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        recs.append((name, describe(*a, **k), e0, e1))
        return out
    return inner

def d_conv(desc, x, w, *rest):
    fl = 2.0 * desc.n * desc.oh * desc.ow * desc.cout * desc.cin * desc.ksize ** 2
    return dict(shape=f"n{desc.n} {desc.h}x{desc.w} {desc.cin}->{desc.cout} k{desc.ksize} s{desc.stride} up{desc.upsample}", flops=fl)

def d_norm(x, *a, **k):
    return dict(shape=str(tuple(x.shape)), bytes=2.0 * x.numel() * x.element_size())

def d_gn(x, *a, **k):
    return dict(shape=str(tuple(x.shape)), bytes=3.0 * x.numel() * x.element_size())

def d_sm(s, dt):
    return dict(shape=str(tuple(s.shape)), bytes=s.numel() * 6.0)

ops.conv2d_tc = wrap("conv_tc", ops.conv2d_tc, d_conv)
ops.conv2d_tc_norm = wrap("conv_tc+norm", ops.conv2d_tc_norm, d_conv)
_gn_inner = wrap("conv_tc+gnst", ops.conv2d_tc_gnstats, d_conv)
def _gnstats(*a, **k):
    n = len(recs)
    out = _gn_inner(*a, **k)
    if out is None:      # layer without a statistics epilogue: nothing was launched
        del recs[n:]
    return out
ops.conv2d_tc_gnstats = _gnstats
ops.conv_out = wrap("conv_out", ops.conv_out, d_conv)
ops.attention = wrap("attention", ops.attention, lambda q, k, vt, n, t: dict(shape=f"n{n} tokens {t} d{vt.shape[1]}", flops=4.0 * n * t * t * vt.shape[1]))
ops.conv2d_direct = wrap("conv_direct", ops.conv2d_direct, d_conv)
ops.rmsnorm_silu = wrap("rmsnorm", ops.rmsnorm_silu, d_norm)
ops.groupnorm_silu = wrap("groupnorm", ops.groupnorm_silu, d_gn)
ops.nchw_to_nhwc = wrap("nchw2nhwc", ops.nchw_to_nhwc, lambda x, cp, dt, *a: dict(shape=str(tuple(x.shape)), bytes=x.numel() * x.element_size() * (1 + cp / x.shape[1])))
ops.reparam = wrap("reparam", ops.reparam, lambda m, *a, **k: dict(shape=str(tuple(m.shape)), bytes=m.numel() * m.element_size() * 2.0))
ops.composite_psnr = wrap("psnr", ops.composite_psnr, lambda r, t, b: dict(shape=str(tuple(r.shape)), bytes=2.0 * r.numel() * r.element_size()))
ops.softmax_rows = wrap("softmax", ops.softmax_rows, d_sm)

torch.manual_seed(0)
model = R.RgbaVAE(R.RgbaAutoencoder(arch).to("cuda", torch.bfloat16))
x = torch.rand(B, 4, S, S, device="cuda").bfloat16()
noise = torch.randn(B, 16, S // 8, S // 8, device="cuda").bfloat16()
for _ in range(2):
    recs.clear()
    recon, _ = model(x, noise=noise)
    ops.composite_psnr(recon, x, [(1.0, 1.0, 1.0)])
torch.cuda.synchronize()
agg = {}
tot = 0.0
for name, d, e0, e1 in recs:
    ms = e0.elapsed_time(e1)
    tot += ms
    key = (name, d["shape"])
    a = agg.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
    a["ms"] += ms; a["n"] += 1; a["flops"] += d.get("flops", 0.0); a["bytes"] += d.get("bytes", 0.0)
print(f"{arch} B={B} {S}x{S}: sum of bracketed launches {tot:.2f} ms")
for (name, shape), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    rate = f"{a['flops']/a['ms']/1e9:8.1f} TFLOP/s" if a["flops"] else f"{a['bytes']/a['ms']/1e6:8.1f} GB/s"
    print(f"{name:12s} {shape:44s} x{a['n']:<3d} {a['ms']:8.3f} ms  {rate}")
