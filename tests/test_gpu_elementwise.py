"""GPU parity of the HBM-bound kernels (through the C ABI) against the CPU oracle / plain torch fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops(lib_built):
    from ragb_vae_b200 import ops as o

    assert torch.cuda.is_available()
    return o


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("c,pixels", [(96, 1000), (192, 77), (384, 513), (96, 1)])
@pytest.mark.parametrize("silu", [True, False])
def test_rmsnorm_silu(ops, dtype, tol, c, pixels, silu):
    g = torch.Generator().manual_seed(c + pixels)
    x = (torch.randn(pixels, c, generator=g) * 3).to(dtype)
    gamma = torch.rand(c, generator=g) + 0.5
    xf = x.float()
    ref = xf / xf.norm(dim=1, keepdim=True).clamp_min(1e-12) * math.sqrt(c) * gamma
    ref = F.silu(ref) if silu else ref
    y = ops.rmsnorm_silu(x.cuda(), gamma.cuda(), silu)
    assert y.dtype == dtype and rel(y, ref) < tol


def test_rmsnorm_zero_pixel_is_finite(ops):
    x = torch.zeros(8, 96)
    y = ops.rmsnorm_silu(x.cuda(), torch.ones(96).cuda(), True)
    assert torch.all(y == 0)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("n,hw,c", [(2, 300, 128), (1, 64, 512), (3, 1025, 256)])
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_silu(ops, dtype, tol, n, hw, c, silu):
    g = torch.Generator().manual_seed(hw + c)
    x = (torch.randn(n, hw, c, generator=g) * 2 + 0.7).to(dtype)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, eps=1e-6).permute(0, 2, 1)
    ref = F.silu(ref) if silu else ref
    y = ops.groupnorm_silu(x.cuda(), gamma.cuda(), beta.cuda(), 32, 1e-6, silu)
    assert rel(y, ref) < tol


@pytest.mark.parametrize("rows,cols", [(5, 1024), (3, 64), (2, 16384), (1, 4)])
def test_softmax_rows(ops, rows, cols):
    g = torch.Generator().manual_seed(rows * cols)
    s = torch.randn(rows, cols, generator=g) * 8
    ref = torch.softmax(s, dim=1)
    assert rel(ops.softmax_rows(s.cuda(), torch.float32), ref) < 2e-6
    assert rel(ops.softmax_rows(s.cuda(), torch.bfloat16), ref) < 5e-3


def test_layout_roundtrip(ops):
    x = torch.randn(2, 4, 6, 10)
    y = ops.nchw_to_nhwc(x.cuda(), 16, torch.float32, 2.0, -1.0)
    assert y.shape == (2, 6, 10, 16)
    assert torch.equal(y[..., :4].cpu(), (x * 2 - 1).permute(0, 2, 3, 1))
    assert torch.all(y[..., 4:] == 0)
    back = ops.nhwc_to_nchw(y, 4, torch.float32)
    assert torch.equal(back.cpu(), x * 2 - 1)
    yb = ops.nchw_to_nhwc(x.cuda().bfloat16(), 4, torch.bfloat16)
    assert torch.equal(yb.cpu(), x.bfloat16().permute(0, 2, 3, 1))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 5e-3)])
def test_reparam_and_kl(ops, dtype, tol):
    g = torch.Generator().manual_seed(5)
    mom = torch.randn(3, 32, 8, 12, generator=g)
    mom[0, 16] = 40.0   # logvar clamps at +20
    mom[1, 17] = -40.0  # and at -30
    mom = mom.to(dtype)
    eps = torch.randn(3, 16, 8, 12, generator=g).to(dtype)
    d = O.DiagonalGaussianDistribution(mom.float())
    z, kl = ops.reparam(mom.cuda(), eps.cuda(), want_kl=True)
    assert rel(z, d.sample(noise=eps.float())) < tol
    assert rel(kl, d.kl()) < 1e-5
    z2, _ = ops.reparam(mom.cuda(), eps.cuda(), z_shift=0.1159, z_scale=0.3611)
    assert rel(z2, (d.sample(noise=eps.float()) - 0.1159) * 0.3611) < tol
    _, kl_only = ops.reparam(mom.cuda(), None, want_kl=True)
    assert rel(kl_only, d.kl()) < 1e-5
    with pytest.raises(ValueError):
        ops.reparam(mom.cuda(), eps[:, :8].cuda())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (1, 17, 13), (3, 96, 128)])
def test_recon_loss(ops, dtype, n, h, w):
    g = torch.Generator().manual_seed(h * w)
    p = (torch.rand(n, 4, h, w, generator=g) * 2 - 1).to(dtype)
    t = (torch.rand(n, 4, h, w, generator=g) * 2 - 1).to(dtype)
    per = ops.recon_loss_per_sample(p.cuda(), t.cuda(), O.EB, O.EB2)
    ref = O.reconstruction_loss(p.float(), t.float(), reduce_mean=False)
    assert per.shape == (n,) and abs(float(per.mean()) - float(ref)) / abs(float(ref)) < 1e-5
    naive = ops.recon_loss_per_sample(p.cuda(), t.cuda(), O.EB, O.EB2, naive_mse=True)
    refn = O.reconstruction_loss(p.float(), t.float(), reduce_mean=False, use_naive_mse=True)
    assert abs(float(naive.mean()) - float(refn)) / float(refn) < 1e-5
    zero = ops.recon_loss_per_sample(p.cuda(), p.cuda(), O.EB, O.EB2)
    assert torch.all(zero == 0)


def test_alpha_vae_loss_module_reduce_rules(ops):
    from ragb_vae_b200 import AlphaVaeLoss

    g = torch.Generator().manual_seed(9)
    p, t = torch.rand(2, 4, 24, 40, generator=g) * 2 - 1, torch.rand(2, 4, 24, 40, generator=g) * 2 - 1
    for rm in (True, False):
        for naive in (True, False):
            got = AlphaVaeLoss(reduce_mean=rm, use_naive_mse=naive).reconstruction_loss(p.cuda(), t.cuda())
            ref = O.reconstruction_loss(p, t, reduce_mean=rm, use_naive_mse=naive)
            assert abs(float(got) - float(ref)) / float(ref) < 1e-5, (rm, naive)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,h,w", [(2, 64, 64), (1, 9, 7)])
def test_composite_psnr(ops, dtype, n, h, w):
    recon = O.synthetic_rgba(n, h, w, seed=3, structured=True).to(dtype)
    target = O.synthetic_rgba(n, h, w, seed=4, structured=True).to(dtype)
    out = ops.composite_psnr(recon.cuda(), target.cuda(), [(1, 1, 1), (0, 0, 0), (0.2, 0.4, 0.6)]).cpu()
    ref = O.validation_metrics(recon.float(), target.float(), backgrounds=(1.0, 0.0, (0.2, 0.4, 0.6)))
    assert torch.allclose(out[:, 0], ref[1.0], atol=2e-4)
    assert torch.allclose(out[:, 1], ref[0.0], atol=2e-4)
    assert torch.allclose(out[:, 2], ref[(0.2, 0.4, 0.6)], atol=2e-4)
    assert torch.allclose(out[:, 3], ref["alpha_mae"], rtol=1e-5)
    same = ops.composite_psnr(recon.cuda(), recon.cuda(), [(1, 1, 1)]).cpu()
    assert torch.allclose(same[:, 0], torch.full((n,), 80.0)) and torch.all(same[:, 1] == 0)  # mse clamps at 1e-8


def test_validation_helpers(ops):
    from ragb_vae_b200 import compute_psnr, validation_metrics

    a, b = torch.rand(2, 3, 16, 16), torch.rand(2, 3, 16, 16)
    assert torch.allclose(compute_psnr(a.cuda(), b.cuda()).cpu(), O.compute_psnr(a, b), atol=2e-4)
    x, y = O.synthetic_rgba(2, 16, 16, 1), O.synthetic_rgba(2, 16, 16, 2)
    m = validation_metrics(x.cuda(), y.cuda())
    ref = O.validation_metrics(x, y)
    assert torch.allclose(m["psnr_white"].cpu(), ref[1.0], atol=2e-4)
    assert torch.allclose(m["psnr_black"].cpu(), ref[0.0], atol=2e-4)
    assert torch.allclose(m["alpha_mae"].cpu(), ref["alpha_mae"], rtol=1e-5)
    with pytest.raises(ValueError):
        validation_metrics(x.cuda(), y.cuda(), backgrounds=("magenta",))


