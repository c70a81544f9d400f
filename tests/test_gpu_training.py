"""GPU parity of the training-step building blocks against torch autograd on the CPU oracle functions."""
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
def T(lib_built):
    from ragb_vae_b200 import training

    assert torch.cuda.is_available()
    return training


@pytest.mark.parametrize("reduce_mean", [True, False])
@pytest.mark.parametrize("naive", [True, False])
def test_recon_loss_backward(T, reduce_mean, naive):
    g = torch.Generator().manual_seed(1)
    p = (torch.rand(2, 4, 20, 28, generator=g) * 2 - 1).requires_grad_(True)
    t = torch.rand(2, 4, 20, 28, generator=g) * 2 - 1
    loss = O.reconstruction_loss(p, t, reduce_mean=reduce_mean, use_naive_mse=naive)
    (3.0 * loss).backward()
    got = T.recon_loss_backward(p.detach().cuda(), t.cuda(), O.EB, O.EB2, reduce_mean, naive, grad_output=3.0)
    assert rel(got, p.grad) < 1e-5
    got_bf = T.recon_loss_backward(p.detach().cuda().bfloat16(), t.cuda().bfloat16(), O.EB, O.EB2, reduce_mean, naive, 3.0)
    assert rel(got_bf.float(), p.grad) < 2e-2


def test_reparam_backward(T):
    g = torch.Generator().manual_seed(2)
    mom = torch.randn(2, 32, 6, 5, generator=g)
    mom[0, 16] = 35.0   # outside the clamp: zero logvar gradient
    mom[1, 17] = -40.0
    mom.requires_grad_(True)
    eps = torch.randn(2, 16, 6, 5, generator=g)
    dz = torch.randn(2, 16, 6, 5, generator=g)
    d = O.DiagonalGaussianDistribution(mom)
    (d.sample(noise=eps) * dz).sum().backward(retain_graph=True)
    got = T.reparam_backward(mom.detach().cuda(), eps.cuda(), dz.cuda())
    assert rel(got, mom.grad) < 1e-5
    mom.grad = None
    ((d.sample(noise=eps) * dz).sum() + 0.3 * d.kl().sum()).backward()
    got = T.reparam_backward(mom.detach().cuda(), eps.cuda(), dz.cuda(), kl_weight=0.3)
    assert rel(got, mom.grad) < 1e-5


@pytest.mark.parametrize("c", [96, 192, 384])
@pytest.mark.parametrize("silu", [True, False])
def test_rmsnorm_silu_backward(T, c, silu):
    g = torch.Generator().manual_seed(c)
    x = (torch.randn(37, c, generator=g) * 2).requires_grad_(True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    dy = torch.randn(37, c, generator=g)
    y = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12) * math.sqrt(c) * gamma
    y = F.silu(y) if silu else y
    (y * dy).sum().backward()
    dx, dg = T.rmsnorm_silu_backward(x.detach().cuda(), gamma.detach().cuda(), dy.cuda(), silu)
    assert rel(dx, x.grad) < 1e-5 and rel(dg, gamma.grad) < 1e-5
    dxb, dgb = T.rmsnorm_silu_backward(x.detach().cuda().bfloat16(), gamma.detach().cuda(), dy.cuda().bfloat16(), silu)
    assert rel(dxb.float(), x.grad) < 2e-2 and rel(dgb, gamma.grad) < 2e-2


@pytest.mark.parametrize("cin,cout,k", [(96, 96, 3), (192, 96, 3), (96, 192, 1)])
def test_conv_dgrad_matches_autograd(T, cin, cout, k):
    g = torch.Generator().manual_seed(cin + cout + k)
    x = torch.randn(2, cin, 12, 136, generator=g).requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).bfloat16().float()
    dy = torch.randn(2, cout, 12, 136, generator=g).bfloat16().float()
    F.conv2d(x, w, padding=k // 2).backward(dy)
    dx = T.conv_dgrad(dy.permute(0, 2, 3, 1).contiguous().cuda().bfloat16(), w.cuda())
    assert rel(dx.float().permute(0, 3, 1, 2), x.grad) < 5e-3


def test_flat_adamw_matches_torch(T):
    torch.manual_seed(0)
    ref_params = [torch.nn.Parameter(torch.randn(257, 33)), torch.nn.Parameter(torch.randn(1000)), torch.nn.Parameter(torch.randn(3, 5, 7))]
    mine = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref_params]
    opt_ref = torch.optim.AdamW(ref_params, lr=1e-2, betas=(0.5, 0.9), eps=1e-8, weight_decay=0.01)
    opt = T.FlatAdamW(mine, lr=1e-2, betas=(0.5, 0.9), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0)
    for step in range(4):
        grads = [torch.randn_like(p) * (5.0 if step == 1 else 0.05) for p in ref_params]   # step 1 exercises the clip
        for p, gr in zip(ref_params, grads):
            p.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        opt_ref.step()
        opt.zero_grad()
        for i, gr in enumerate(grads):
            opt.grad_view(i).copy_(gr.cuda() * 2.0)      # as if summed over 2 ranks
        opt.step(grad_scale=0.5)
        for p, q in zip(ref_params, mine):
            assert rel(q, p) < 1e-5, step
    # bf16 model parameters alias a flat bf16 buffer that tracks the fp32 master
    pb = [torch.nn.Parameter(torch.randn(64, 64).cuda().bfloat16())]
    ob = T.FlatAdamW(pb, lr=1e-1, max_grad_norm=None)
    ob.grad_view(0).fill_(1.0)
    before = pb[0].detach().clone()
    ob.step()
    assert pb[0].dtype == torch.bfloat16 and not torch.equal(pb[0], before)
    assert torch.equal(pb[0].detach().reshape(-1), ob.master.bfloat16())


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(1, 8, 64, 64, 64, 3), (2, 12, 136, 96, 96, 3), (1, 16, 70, 192, 96, 3),
                                              (2, 9, 64, 96, 192, 1), (1, 8, 128, 384, 384, 3),
                                              # Cin tiles of 128: one N = 192 MMA per 64-channel block covers the three taps
                                              (2, 10, 72, 128, 128, 3), (1, 6, 64, 256, 512, 3), (1, 7, 200, 512, 256, 3)])
def test_conv_wgrad_matches_autograd(T, n, h, w, cin, cout, k):
    g = torch.Generator().manual_seed(n + h + w + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = torch.zeros(cout, cin, k, k, requires_grad=True)
    bias = torch.zeros(cout, requires_grad=True)
    dy = torch.randn(n, cout, h, w, generator=g).bfloat16().float()
    F.conv2d(x, wt, bias, padding=k // 2).backward(dy)
    dw, db = T.conv_wgrad(x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16(), dy.permute(0, 2, 3, 1).contiguous().cuda().bfloat16(), k)
    assert dw.shape == wt.shape
    assert rel(dw, wt.grad) < 2e-3
    assert rel(db, bias.grad) < 1e-4
