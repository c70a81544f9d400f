"""GPU parity of the two convolution kernels (C ABI) against torch.nn.functional.conv2d on the CPU:
rv_conv2d_direct (fp32 CUDA cores) and rv_conv2d_tc (tcgen05 + TMA, bf16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from ragb_vae_b200._lib import RV_BF16, RV_F32  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops(lib_built):
    from ragb_vae_b200 import ops as o

    assert torch.cuda.is_available()
    return o


def ref_conv(x, w, b, k, stride, upsample):
    """x NCHW fp32; the three conv flavours of the VAE."""
    if upsample:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    if k == 3 and stride == 2:
        return F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    return F.conv2d(x, w, b, padding=k // 2)


CASES = [  # n, h, w, cin, cout, k, stride, upsample
    (2, 16, 16, 64, 64, 3, 1, False),
    (1, 24, 40, 96, 96, 3, 1, False),      # BK=32 path, non-pow2 width
    (1, 8, 8, 16, 384, 3, 1, False),       # BK=16 path (decoder conv_in), two N tiles
    (2, 16, 24, 128, 128, 3, 2, False),    # stride-2 parity view
    (1, 12, 20, 192, 96, 3, 1, True),      # fused nearest-x2 upsample, phase-folded weights
    (1, 16, 16, 96, 192, 1, 1, False),     # 1x1 shortcut
    (1, 32, 32, 96, 4, 3, 1, False),       # conv_out: Cout=4 padded to one 16-wide tile
    (1, 128, 130, 64, 32, 3, 1, False),    # 128-wide tiles with a ragged edge
    (3, 8, 8, 512, 512, 3, 1, False),      # two 256-wide N tiles, deep K
    (2, 20, 200, 96, 96, 3, 1, False),     # halo-reuse kernel: two column blocks (ragged), SW128 + SW64 K blocks
    (1, 9, 130, 96, 96, 3, 1, False),      # halo-reuse kernel: odd height (masked last row), 2-pixel second block
    (1, 12, 256, 64, 128, 3, 1, False),    # halo-reuse kernel: Cin = 64 (no SW64 block), Cout = 128
    (2, 70, 128, 96, 32, 3, 1, False),     # halo-reuse kernel: several strips per column, narrow Cout
    (1, 16, 192, 96, 4, 3, 1, False),      # halo-reuse kernel as conv_out: Cout = 4, generic (NCHW / clamp) epilogue
]


@pytest.mark.parametrize("case", CASES)
def test_direct_conv_fp32(ops, case):
    n, h, w, cin, cout, k, stride, up = case
    g = torch.Generator().manual_seed(sum(case[:5]))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = ref_conv(x, wt, b, k, stride, up)
    wp = ops.pack_conv_weights_direct(wt.cuda())
    oh, ow = ref.shape[2:]
    # NHWC in / NHWC out
    xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = torch.empty(n, oh, ow, cout, device="cuda")
    d = ops.make_desc(n, h, w, cin, cout, k, stride, up, x_dtype=RV_F32, y_dtype=RV_F32)
    ops.conv2d_direct(d, xh, wp, b.cuda(), None, y)
    assert rel(y.permute(0, 3, 1, 2), ref) < 2e-6
    # NCHW in / NCHW out with residual, input affine, output affine + clamp
    res = torch.randn(n, cout, oh, ow, generator=g)
    y2 = torch.empty(n, cout, oh, ow, device="cuda")
    d2 = ops.make_desc(n, h, w, cin, cout, k, stride, up, x_dtype=RV_F32, y_dtype=RV_F32, x_nchw=True, y_nchw=True,
                       in_scale=2.0, in_shift=-1.0, out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0))
    ops.conv2d_direct(d2, x.cuda(), wp, b.cuda(), res.cuda(), y2)
    ref2 = torch.clamp((ref_conv(x * 2 - 1, wt, b, k, stride, up) + res) * 0.5 + 0.5, 0, 1)
    assert rel(y2, ref2) < 2e-6


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_bf16(ops, case):
    n, h, w, cin, cout, k, stride, up = case
    g = torch.Generator().manual_seed(sum(case[:5]) + 1)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    ref = ref_conv(x.float(), wt, b, k, stride, up)
    oh, ow = ref.shape[2:]
    wp = ops.pack_conv_weights_tc(wt.cuda(), up)
    xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    res = torch.randn(n, oh, ow, cout, generator=g).bfloat16()
    # bf16 NHWC out with residual
    y = torch.empty(n, oh, ow, cout, dtype=torch.bfloat16, device="cuda")
    d = ops.make_desc(n, h, w, cin, cout, k, stride, up, x_dtype=RV_BF16, y_dtype=RV_BF16)
    ops.conv2d_tc(d, xh, wp, wp.shape[1], b.cuda(), res.cuda(), y)
    tol = 8e-3 if up else 4e-3  # folded phase weights are re-rounded to bf16
    assert rel(y.float().permute(0, 3, 1, 2), ref + res.float().permute(0, 3, 1, 2)) < tol
    # fp32 NCHW out, affine + clamp epilogue
    y2 = torch.empty(n, cout, oh, ow, dtype=torch.float32, device="cuda")
    d2 = ops.make_desc(n, h, w, cin, cout, k, stride, up, x_dtype=RV_BF16, y_dtype=RV_F32, y_nchw=True, out_scale=0.5,
                       out_shift=0.5, clamp=(0.0, 1.0))
    ops.conv2d_tc(d2, xh, wp, wp.shape[1], b.cuda(), None, y2)
    assert rel(y2, torch.clamp(ref * 0.5 + 0.5, 0, 1)) < tol


def test_tc_gemm_modes(ops):
    """The attention GEMMs: strided operands, alpha scaling, fp32 output, per-row bias."""
    g = torch.Generator().manual_seed(11)
    t, c = 256, 384
    qk = torch.randn(t, 2 * c, generator=g).bfloat16()
    s = torch.empty(t, t, dtype=torch.float32, device="cuda")
    qkc = qk.cuda()
    d = ops.make_desc(1, 1, t, c, t, 1, 1, False, x_dtype=RV_BF16, y_dtype=RV_F32, x_cstride=2 * c, y_cstride=t,
                      bias_mode=0, alpha=0.05)
    ops.conv2d_tc(d, qkc, qkc[:, c:], 2 * c, None, None, s)
    ref = 0.05 * qk[:, :c].float() @ qk[:, c:].float().t()
    assert rel(s, ref) < 1e-5
    wv = torch.randn(c, c, generator=g).bfloat16()
    bv = torch.randn(c, generator=g)
    xn = torch.randn(t, c, generator=g).bfloat16()
    vt = torch.empty(c, t, dtype=torch.bfloat16, device="cuda")
    d2 = ops.make_desc(1, 1, c, c, t, 1, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16, x_cstride=c, y_cstride=t, bias_mode=2)
    ops.conv2d_tc(d2, wv.cuda(), xn.cuda(), c, bv.cuda(), None, vt)
    assert rel(vt.float(), wv.float() @ xn.float().t() + bv[:, None]) < 4e-3


def test_tc_rejects_bad_arguments(ops):
    from ragb_vae_b200._lib import RvError

    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(16, 9 * 24, dtype=torch.bfloat16, device="cuda")
    y = torch.zeros(1, 8, 8, 16, dtype=torch.bfloat16, device="cuda")
    d = ops.make_desc(1, 8, 8, 24, 16, 3, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16, bias_mode=0)
    with pytest.raises(RvError):
        ops.conv2d_tc(d, x, w, 9 * 24, None, None, y)  # cin not a multiple of 16
    d.oh = 7
    with pytest.raises(RvError):
        ops.conv2d_direct(d, x, w.float(), None, None, y)  # inconsistent output size


@pytest.mark.parametrize("w", [40, 136])  # 136 columns: the halo-reuse kernel takes the Cout <= 128 cases
@pytest.mark.parametrize("cout,want_raw,silu", [(96, True, True), (192, False, True), (96, True, False)])
def test_tc_conv_fused_rmsnorm_epilogue(ops, cout, want_raw, silu, w):
    """rv_conv2d_tc_norm: the consumer's QwenImageRMS_norm (+SiLU) applied in the conv epilogue."""
    import math

    g = torch.Generator().manual_seed(cout + int(want_raw))
    n, h, cin = 2, 12, 96
    x = torch.randn(n, cin, h, w, generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    res = torch.randn(n, h, w, cout, generator=g).bfloat16()
    gamma = torch.rand(cout, generator=g) + 0.5
    raw_ref = (ref_conv(x.float(), wt, b, 3, 1, False) + res.float().permute(0, 3, 1, 2))
    nrm = raw_ref / raw_ref.norm(dim=1, keepdim=True).clamp_min(1e-12) * math.sqrt(cout) * gamma.view(1, -1, 1, 1)
    act_ref = F.silu(nrm) if silu else nrm
    wp = ops.pack_conv_weights_tc(wt.cuda())
    xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda") if want_raw else None
    act = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    d = ops.make_desc(n, h, w, cin, cout, 3, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16)
    ops.conv2d_tc_norm(d, xh, wp, wp.shape[1], b.cuda(), res.cuda(), y, act, (gamma * math.sqrt(cout)).cuda(), silu)
    assert rel(act.float().permute(0, 3, 1, 2), act_ref) < 5e-3
    if want_raw:
        assert rel(y.float().permute(0, 3, 1, 2), raw_ref) < 4e-3
    # more than 256 output channels cannot be fused (two accumulator tiles)
    from ragb_vae_b200._lib import RvError

    d2 = ops.make_desc(1, 8, 8, 96, 384, 3, 1, False, x_dtype=RV_BF16, y_dtype=RV_BF16)
    w2 = torch.zeros(384, 9 * 96, dtype=torch.bfloat16, device="cuda")
    y2 = torch.empty(1, 8, 8, 384, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RvError):
        ops.conv2d_tc_norm(d2, xh[:1, :8, :8].contiguous(), w2, 9 * 96, torch.zeros(384, device="cuda"), None, y2, y2.clone(),
                           torch.ones(384, device="cuda"), True)


@pytest.mark.parametrize("d", [384, 512])  # Qwen-Image mid block; Flux mid block (two output passes inside the kernel)
@pytest.mark.parametrize("n_img,tokens", [(1, 128), (2, 512), (1, 2048)])
@pytest.mark.parametrize("scale_up", [1.0, 6.0])  # larger scores exercise the lazy-rescale path
def test_fused_attention(ops, n_img, tokens, scale_up, d):
    g = torch.Generator().manual_seed(tokens + n_img)
    qk = (torch.randn(n_img * tokens, 2 * d, generator=g) * scale_up).bfloat16()
    # make the row maxima grow along the key axis so that several rescales happen
    qk[:, d:] *= torch.linspace(0.2, 1.5, n_img * tokens).view(-1, 1).bfloat16()
    v = torch.randn(n_img, tokens, d, generator=g).bfloat16()
    vt = v.transpose(1, 2).contiguous()
    q = qk[:, :d].float().view(n_img, tokens, d)
    k = qk[:, d:].float().view(n_img, tokens, d)
    ref = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.float().unsqueeze(1)).squeeze(1)
    qkc = qk.cuda()
    out = ops.attention(qkc[:, :d], qkc[:, d:], vt.cuda(), n_img, tokens)
    assert out.shape == (n_img * tokens, d)
    assert rel(out.float().view(n_img, tokens, d), ref) < 1e-2


@pytest.mark.parametrize("d", [384, 512])
@pytest.mark.parametrize("tokens", [256, 640, 1152])  # 640: N tile of 160 columns -> generic epilogue; else the lean one
def test_attention_lse_and_fused_score_gemms(ops, tokens, d):
    """rv_attention_lse's log-sum-exp, and the two score-matrix GEMMs of the attention backward with their elementwise step
    in the epilogue (rv_gemm_rowstat): P = exp2(s - lse) equals softmax(QK^T/sqrt(d)), dS = P * (dP - delta) / sqrt(d) with
    delta = rowdot(dO, O) equals torch autograd's gradient of the scores."""
    g = torch.Generator().manual_seed(tokens + d)
    q = torch.randn(tokens, d, generator=g).bfloat16()
    k = (torch.randn(tokens, d, generator=g) * torch.linspace(0.3, 1.4, tokens).view(-1, 1)).bfloat16()
    v = torch.randn(tokens, d, generator=g).bfloat16()
    d_o = torch.randn(tokens, d, generator=g).bfloat16()
    scale = 1.0 / d ** 0.5
    s_ref = (q.float() @ k.float().t()) * scale
    p_ref = torch.softmax(s_ref, -1)
    o_ref = p_ref @ v.float()
    qc, kc, vc, doc = q.cuda(), k.cuda(), v.cuda(), d_o.cuda()
    out, lse = ops.attention(qc, kc, vc.t().contiguous().view(1, d, tokens), 1, tokens, return_lse=True)
    lse_ref = torch.logsumexp(s_ref, -1) * 1.4426950408889634
    assert float((lse.cpu() - lse_ref).abs().max()) < 2e-3
    assert rel(out.float().cpu(), o_ref) < 1e-2
    p = ops.gemm_rowstat(qc, kc, lse, 1, scale * 1.4426950408889634)
    assert p.dtype == torch.bfloat16 and rel(p.float().cpu(), p_ref) < 6e-3
    assert float((p.float().sum(-1) - 1).abs().max()) < 1e-2
    delta = ops.rowdot(doc, out, scale)
    delta_ref = (d_o.float() * o_ref).sum(-1) * scale
    assert float((delta.cpu() - delta_ref).abs().max()) < 2e-2 * float(delta_ref.abs().max()) + 1e-3
    ds = ops.gemm_rowstat(doc, vc, delta, 2, scale, mul_in=p)
    dp_ref = d_o.float() @ v.float().t()
    ds_ref = p_ref * (dp_ref - (dp_ref * p_ref).sum(-1, keepdim=True)) * scale
    assert rel(ds.float().cpu(), ds_ref) < 2e-2
    with pytest.raises(Exception):
        ops.gemm_rowstat(qc, kc, lse, 2, scale)          # mode 2 without its multiplicand
    with pytest.raises(Exception):
        ops.gemm_rowstat(qc, kc, lse, 1, scale, mul_in=p)  # mode 1 takes none


@pytest.mark.parametrize("shape", [(2, 4, 40, 72), (1, 4, 16, 200), (1, 3, 24, 24)])
def test_hpack_stem_matches_conv2d(ops, shape):
    """conv_in through the horizontally packed loader + 3 vertical taps == the plain 3x3 conv (incl. image borders)."""
    import torch.nn.functional as F
    import ragb_vae_b200 as R

    g = torch.Generator().manual_seed(7)
    n, c, h, w = shape
    x = torch.rand(shape, generator=g)
    xb = x.to(torch.bfloat16).float()
    packed = ops.nchw_to_nhwc_hpack(x.cuda().to(torch.bfloat16), 2.0, -1.0)
    want = torch.zeros(n, h, w, 16)
    xp = F.pad(xb * 2 - 1, (1, 1))
    for dx in range(3):
        want[..., dx * c:(dx + 1) * c] = xp[:, :, :, dx:dx + w].permute(0, 2, 3, 1)
    want[:, :, 0, :c] = 0
    want[:, :, -1, 2 * c:3 * c] = 0
    assert torch.allclose(packed.float().cpu(), want.to(torch.bfloat16).float(), atol=1e-2)
    m = R.RgbaAutoencoder("qwen", in_channels=c).to("cuda", torch.bfloat16)
    conv = m.encoder.conv_in
    ref = F.conv2d((xb * 2 - 1).to(torch.bfloat16).float(), conv.weight2d().float().cpu(), conv.bias.float().cpu(), padding=1)
    outs = []
    for flag in (True, False):
        m.hpack_stem = flag
        st = m._stem(x.cuda().to(torch.bfloat16), conv, 2.0, -1.0)
        outs.append(st.raw.float().permute(0, 3, 1, 2).cpu())
    for o in outs:
        assert float((o - ref).norm() / ref.norm()) < 1e-2


@pytest.mark.parametrize("cin,cout", [(96, 4), (128, 4), (64, 4), (96, 3), (96, 5)])
@pytest.mark.parametrize("shape,f32", [((2, 40, 200), False), ((1, 133, 64), True), ((1, 1024, 1024), False)])
def test_conv_out_kernel_matches_conv2d(ops, cin, cout, shape, f32):
    """rv_conv_out (kernel rows in the MMA's N, weights stationary, dy summed from TMEM by the pixel's thread) against
    F.conv2d on the same bf16-rounded operands, incl. ragged widths / heights, image borders, several strips per column,
    the fused affine + clamp of RgbaVAE.forward and fp32 / bf16 NCHW outputs."""
    n, h, w = shape
    if h * w >= 1 << 20 and (cin, cout) != (96, 4):
        pytest.skip("full resolution once")
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout + h)
    x = torch.randn(n, h, w, cin, generator=g, device="cuda").bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3.0 * cin ** 0.5)).bfloat16().float()
    b = torch.randn(cout, generator=g, device="cuda") * 0.1
    y = torch.empty(n, cout, h, w, dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
    d = ops.make_desc(n, h, w, cin, cout, 3, 1, False, x_dtype=RV_BF16, y_dtype=RV_F32 if f32 else RV_BF16, y_nchw=True,
                      out_scale=0.5, out_shift=0.5, clamp=(0.0, 1.0))
    assert ops.conv_out_eligible(n, h, w, cin, cin, cout, 3, 1, False)
    ops.conv_out(d, x, ops.pack_conv_out_weights(wt), b, y)
    ref = torch.clamp(F.conv2d(x.float().permute(0, 3, 1, 2), wt, b, padding=1) * 0.5 + 0.5, 0.0, 1.0)
    err = float((y.float() - ref).abs().max())
    assert err < (2e-4 if f32 else 5e-3), err
    # borders and the strip seams specifically (the first / last rows and columns are where halo handling lives)
    for sl in ((slice(None), slice(None), slice(0, 2)), (slice(None), slice(None), slice(-2, None)),
               (slice(None), slice(None), slice(None), slice(0, 2)), (slice(None), slice(None), slice(None), slice(-2, None))):
        assert float((y.float()[sl] - ref[sl]).abs().max()) < (2e-4 if f32 else 5e-3)


@pytest.mark.parametrize("n,h,w,cin,cout,res", [(2, 64, 64, 256, 256, True), (1, 32, 96, 512, 512, False), (3, 48, 40, 128, 256, False),
                                               (2, 32, 32, 128, 512, True), (1, 40, 24, 512, 256, False)])
def test_conv_gnstats_epilogue(ops, n, h, w, cin, cout, res):
    _gnstats_case(ops, n, h, w, cin, cout, res, False)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 24, 40, 256, 256), (1, 16, 16, 512, 512)])
def test_conv_gnstats_epilogue_upsample(ops, n, h, w, cin, cout):
    """The nearest-x2 + 3x3 conv (four phase launches): one segment of partials per phase."""
    _gnstats_case(ops, n, h, w, cin, cout, False, True)


def _gnstats_case(ops, n, h, w, cin, cout, res, up):
    """rv_conv2d_tc_gnstats: same output bits as rv_conv2d_tc, and statistics equal to rv_groupnorm_stats of that output
    (the fused ones are taken before the bf16 rounding: relative 2e-3 on the sums of squares); ragged tiles, odd tile counts,
    one and two N tiles, with and without residual; batch independence of a sample's statistics."""
    import ctypes as C
    from ragb_vae_b200 import _lib
    g = torch.Generator().manual_seed(n * 1000 + cout)
    x = torch.randn(n, h, w, cin, generator=g).bfloat16().cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    s_ = 2 if up else 1
    r = torch.randn(n, h * s_, w * s_, cout, generator=g).bfloat16().cuda() if res else None
    wp = ops.pack_conv_weights_tc(wt, up)
    desc = ops.make_desc(n, h, w, cin, cout, 3, 1, up, x_dtype=_lib.RV_BF16, y_dtype=_lib.RV_BF16)
    y0 = torch.empty(n, h * s_, w * s_, cout, dtype=torch.bfloat16, device="cuda")
    ops.conv2d_tc(desc, x, wp, wp.shape[1], bias, r, y0)
    y1 = torch.empty_like(y0)
    stats = ops.conv2d_tc_gnstats(desc, x, wp, wp.shape[1], bias, r, y1, 32)
    assert stats is not None and tuple(stats.shape) == (n, 32, 2)
    d128 = ops.make_desc(n, h, w, cin, 128, 3, 1, up, x_dtype=_lib.RV_BF16, y_dtype=_lib.RV_BF16)
    assert ops.conv2d_tc_gnstats(d128, x, wp[:128], wp.shape[1], bias[:128].contiguous(), None, y1[..., :128].contiguous(), 32) is None
    assert torch.equal(y0, y1)
    ref = torch.empty((n, 32, 2), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().rv_groupnorm_stats(ops._ptr(y0), ops._ptr(ref), n, h * w * s_ * s_, cout, 32, _lib.RV_BF16, ops._stream(y0)), "stats")
    cnt = h * w * s_ * s_ * (cout // 32)
    mean_f, mean_r = stats[..., 0] / cnt, ref[..., 0] / cnt
    var_f, var_r = stats[..., 1] / cnt - mean_f ** 2, ref[..., 1] / cnt - mean_r ** 2
    assert float((mean_f - mean_r).abs().max()) < 2e-3 * float(var_r.sqrt().max())
    assert float(((var_f - var_r).abs() / var_r).max()) < 4e-3
    # a sample's statistics do not depend on the batch it is in
    if n > 1:
        d1 = ops.make_desc(1, h, w, cin, cout, 3, 1, up, x_dtype=_lib.RV_BF16, y_dtype=_lib.RV_BF16)
        ya = torch.empty(1, h * s_, w * s_, cout, dtype=torch.bfloat16, device="cuda")
        sa = ops.conv2d_tc_gnstats(d1, x[-1:].contiguous(), wp, wp.shape[1], bias, None if r is None else r[-1:].contiguous(), ya, 32)
        if sa is not None:
            assert torch.equal(sa[0], stats[-1])
