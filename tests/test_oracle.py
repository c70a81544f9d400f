"""CPU tests of the oracle itself: pinned against the committed golden vectors (made by
scripts/make_golden.py, which also checked the flux oracle against the independent BFL/torchtitan
auto-encoder), exact public parameter counts and algebraic identities."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("arch,n3,n4", [("flux", 83_819_683, 83_821_988), ("qwen", 126_892_531, 126_897_716)])
def test_param_counts_match_public_models(arch, n3, n4):
    assert sum(p.numel() for p in O.OracleVAE(arch, 3, 3).parameters()) == n3
    assert sum(p.numel() for p in O.OracleVAE(arch, 4, 4).parameters()) == n4


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_oracle_reproduces_golden(arch, golden, oracle_model):
    g = golden(arch)
    vae = oracle_model(arch)
    wsum = sum(p.double().abs().sum() for p in vae.parameters()).float()
    assert torch.allclose(wsum, g["weight_checksum"][0], rtol=1e-6), "seed-0 init differs from the fixture's"
    x = O.synthetic_rgba(1, 256, 256, seed=1)
    noise = torch.randn(1, 16, 32, 32, generator=torch.Generator().manual_seed(2))
    assert torch.allclose(x.double().sum().float(), g["x_checksum"][0], rtol=1e-6)
    recon, post, z = O.rgba_vae_forward(vae, x, noise)
    assert rel(post.parameters, g["moments"]) < 1e-5
    assert rel(z, g["z"]) < 1e-5
    assert rel(recon, g["recon"]) < 1e-5
    m = O.validation_metrics(recon, x)
    assert abs(float(m[1.0]) - float(g["psnr_white"])) < 1e-3
    assert abs(float(m[0.0]) - float(g["psnr_black"])) < 1e-3


def test_flux_oracle_matches_independent_bfl_implementation(golden):
    g = golden("flux")
    assert rel(g["titan_moments"], g["moments"]) < 1e-5
    assert rel(g["titan_decoded"], g["decoded"]) < 1e-5
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        man = json.load(f)
    flux = [r for r in man["runs"] if r["arch"] == "flux"][0]
    assert flux["titan_vs_oracle_moments_maxabs"] < 1e-4 and flux["titan_vs_oracle_decoded_maxabs"] < 1e-4


def test_causal_conv3d_single_frame_equals_conv2d_last_tap():
    torch.manual_seed(0)
    c = O.QwenCausalConv3d(5, 7, 3, padding=1)
    x = torch.randn(2, 5, 9, 11)
    lit = c(x.unsqueeze(2)).squeeze(2)
    assert torch.allclose(lit, c.forward_frame(x), atol=1e-5)
    c1 = O.QwenCausalConv3d(5, 7, 1)
    assert torch.allclose(c1(x.unsqueeze(2)).squeeze(2), c1.forward_frame(x), atol=1e-5)


def test_rms_norm_matches_definition():
    n = O.QwenRMSNorm(12, images=False)
    with torch.no_grad():
        n.gamma.copy_(torch.rand_like(n.gamma) + 0.5)
    x = torch.randn(2, 12, 4, 5)
    ref = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12) * 12 ** 0.5 * n.gamma.reshape(1, -1, 1, 1)
    assert torch.allclose(n(x), ref, atol=1e-6)


def test_posterior_and_losses_small():
    torch.manual_seed(3)
    mom = torch.randn(2, 8, 3, 3)
    mom[0, 4] = 50.0  # clamp hi
    mom[1, 5] = -50.0  # clamp lo
    d = O.DiagonalGaussianDistribution(mom)
    eps = torch.randn(2, 4, 3, 3)
    assert torch.allclose(d.sample(noise=eps), d.mean + torch.exp(0.5 * d.logvar) * eps)
    assert d.logvar.max() <= 20 and d.logvar.min() >= -30
    kl = 0.5 * (d.mean ** 2 + d.var - 1 - d.logvar).sum(dim=(1, 2, 3))
    assert torch.allclose(d.kl(), kl)
    p, t = torch.rand(2, 4, 6, 6) * 2 - 1, torch.rand(2, 4, 6, 6) * 2 - 1
    # reduce rules (losses.py:117-123)
    m = O.reconstruction_loss(p, t, reduce_mean=True)
    s = O.reconstruction_loss(p, t, reduce_mean=False)
    assert torch.allclose(m * 3 * 36, s, rtol=1e-5)
    # identical alpha -> plain premultiplied MSE
    t2 = t.clone(); t2[:, 3] = p[:, 3]
    a = (p[:, 3:] + 1) / 2
    assert torch.allclose(O.reconstruction_loss(p, t2, True), ((t2[:, :3] * a - p[:, :3] * a) ** 2).mean(), atol=1e-6)


def test_composite_errors_and_triplet():
    x = torch.rand(2, 4, 5, 5)
    assert torch.allclose(O.composite_over_background(x, 1.0), x[:, :3] * x[:, 3:] + (1 - x[:, 3:]))
    assert torch.allclose(O.composite_over_background(x, (0.2, 0.4, 0.6))[:, 1],
                          x[:, 1] * x[:, 3] + 0.4 * (1 - x[:, 3]))
    with pytest.raises(ValueError):
        O.composite_over_background(x, (1.0, 0.0))
    with pytest.raises(ValueError):
        O.composite_over_background(x, torch.zeros(3, 4, 4))
    t = x * 2 - 1
    tri = O.build_detail_augmented_triplet(t)
    assert tri.shape[0] == 6 and torch.all(tri[2:, 3] == 1.0)
    with pytest.raises(ValueError):
        O.build_detail_augmented_triplet(t[:, :3])
    d = O.DiagonalGaussianDistribution(torch.randn(6, 8, 2, 2))
    a, b, c = O.split_triplet_distribution(d)
    assert a.parameters.shape[0] == 2


def test_adapt_matches_reference_rule():
    for arch in ("flux", "qwen"):
        m = O.OracleVAE(arch, 3, 3)
        w_in, w_out = m.encoder.conv_in.weight.clone(), m.decoder.conv_out.weight.clone()
        O.adapt_vae_to_rgba(m, alpha_bias_init=0.7)
        assert m.encoder.conv_in.weight.shape[1] == 4 and m.decoder.conv_out.weight.shape[0] == 4
        assert torch.equal(m.encoder.conv_in.weight[:, :3], w_in) and torch.all(m.encoder.conv_in.weight[:, 3] == 0)
        assert torch.equal(m.decoder.conv_out.weight[:3], w_out) and torch.all(m.decoder.conv_out.weight[3] == 0)
        assert float(m.decoder.conv_out.bias[3]) == pytest.approx(0.7)
        assert m.config.in_channels == 4 and m.config.out_channels == 4


def test_training_step_oracle_invariants():
    """oracle.training_step (rgba_vae_stage.py:433-518): what autograd must give on one frame -- no gradient for the
    video-only temporal convs, exact zeros in the unused temporal taps of the causal 3-D kernels, finite everything, and
    a reference-KL term that only appears with a reference VAE and a positive scale."""
    import copy

    import torch

    from oracle import vae_oracle as O

    vae = copy.deepcopy(O.build_oracle("qwen", seed=0))
    x = O.synthetic_rgba(1, 32, 32, seed=3)
    noise = torch.randn(1, 16, 4, 4, generator=torch.Generator().manual_seed(4))
    metrics, grads = O.training_step(vae, x, noise, kl_scale=1e-6)
    assert set(metrics) == {"train/recon", "train/kl", "train/loss"}
    assert all(torch.isfinite(v).all() for v in grads.values())
    names = dict(vae.named_parameters())
    assert not any(".time_conv." in n for n in grads) and any(".time_conv." in n for n in names)
    w = grads["encoder.down_blocks.0.conv1.weight"]
    assert w.dim() == 5 and float(w[:, :, :-1].abs().max()) == 0.0 and float(w[:, :, -1].abs().max()) > 0.0
    assert abs(float(metrics["train/loss"]) - float(metrics["train/recon"]) - 1e-6 * float(metrics["train/kl"])) < 1e-6 * abs(
        float(metrics["train/loss"])) + 1e-9
    ref = copy.deepcopy(O.build_oracle("qwen", seed=5))
    m2, g2 = O.training_step(vae, x, noise, kl_scale=1e-6, ref_vae=ref, ref_kl_scale=0.5)
    assert "train/ref_kl" in m2 and float(m2["train/ref_kl"]) > 0.0
    assert float((g2["encoder.conv_in.weight"] - grads["encoder.conv_in.weight"]).abs().max()) > 0.0
    m3, _ = O.training_step(vae, x, noise, kl_scale=1e-6, ref_vae=ref, ref_kl_scale=0.0)
    assert "train/ref_kl" not in m3


def test_background_blend_and_batch_assembly_oracle():
    import torch

    from oracle import vae_oracle as O

    t = torch.rand(4, 6, 5, generator=torch.Generator().manual_seed(1))
    c = torch.tensor([0.3, 0.5, 0.9])
    out = O.background_blend(t, c)
    assert out.shape == (4, 6, 5) and float(out[3].min()) == 1.0
    a = t[3:4]
    assert torch.allclose(out[:3], t[:3] * a + c.view(3, 1, 1) * (1 - a))
    b = {"component": torch.zeros(2, 4, 3, 3), "composite": torch.ones(2, 4, 3, 3), "background": torch.full((2, 4, 3, 3), 0.5)}
    assert O.build_training_batch(b).shape[0] == 4
    assert O.build_training_batch(b, torch.tensor([True, False])).shape[0] == 5
    with pytest.raises(ValueError):
        O.build_training_batch({"component": b["component"]})
