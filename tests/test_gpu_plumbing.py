"""GPU parity of the data-format kernels either side of the VAE against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def P(lib_built):
    from ragb_vae_b200 import plumbing

    assert torch.cuda.is_available()
    return plumbing


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_triplet_augmentation_and_split(P, dtype, tol):
    from ragb_vae_b200 import DiagonalGaussianDistribution

    t = (O.synthetic_rgba(3, 24, 40, seed=5, structured=True) * 2 - 1)
    ref = O.build_detail_augmented_triplet(t)
    got = P.build_detail_augmented_triplet(t.cuda().to(dtype))
    assert got.shape == (9, 4, 24, 40) and got.dtype == dtype
    assert torch.allclose(got.float().cpu(), ref, atol=tol)
    assert torch.all(got[3:, 3] == 1.0)
    assert torch.equal(got[:3].float().cpu(), t.to(dtype).float())
    with pytest.raises(ValueError):
        P.build_detail_augmented_triplet(t[:, :3].cuda())
    post = DiagonalGaussianDistribution(torch.randn(6, 32, 4, 4, device="cuda"))
    a, b, c = P.split_triplet_distribution(post)
    assert a.parameters.shape[0] == 2 and torch.equal(b.parameters, post.parameters[2:4])
    with pytest.raises(ValueError):
        P.split_triplet_distribution(DiagonalGaussianDistribution(torch.randn(4, 32, 4, 4, device="cuda")))


def test_pack_unpack_latents(P):
    z = torch.randn(2, 16, 12, 20)
    ref = O.pack_latents(z)
    got = P.pack_latents(z.cuda())
    assert torch.equal(got.cpu(), ref)
    back = P.unpack_latents(got, 96, 160)
    assert torch.equal(back.cpu(), z) and torch.equal(O.unpack_latents(ref, 96, 160), z)
    # fused Flux normalisation on the way in and out (flux_kontext_textalpha.py:330-332,497)
    zn = P.pack_latents(z.cuda(), shift=0.1159, scale=0.3611)
    assert torch.allclose(zn.cpu(), O.pack_latents((z - 0.1159) * 0.3611), atol=1e-6)
    rt = P.unpack_latents(zn, 96, 160, shift=0.1159, scale=1.0 / 0.3611)
    assert torch.allclose(rt.cpu(), z, atol=1e-5)
    with pytest.raises(ValueError):
        P.unpack_latents(got, 96, 96)
    from ragb_vae_b200._lib import RvError

    with pytest.raises(RvError):
        P.pack_latents(torch.randn(1, 16, 5, 6).cuda())


def test_uint8_ingest_egress(P):
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 9, 7, 4), generator=g, dtype=torch.uint8)
    t = P.rgba_u8_to_tensor(img.cuda(), torch.float32)
    ref = torch.stack([O.load_rgba_u8(i) for i in img])
    assert torch.equal(t.cpu(), ref)
    tv = P.rgba_u8_to_tensor(img.cuda(), torch.float32, vae_range=True)
    assert torch.allclose(tv.cpu(), ref * 2 - 1, atol=1e-6)
    x = torch.rand(2, 4, 9, 7, generator=g) * 1.4 - 0.2
    out = P.tensor_to_rgba_u8(x.cuda())
    assert torch.equal(out.cpu(), torch.stack([O.save_rgba_u8(i) for i in x]))
    # uint8 -> tensor -> uint8 is the identity
    assert torch.equal(P.tensor_to_rgba_u8(t).cpu(), img)
    with pytest.raises(ValueError):
        P.rgba_u8_to_tensor(img[..., :3].cuda())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_random_background_blend_matches_reference_blend(P, dtype):
    """Batched RandomBackgroundBlend vs the per-sample reference formula with the same mask and colours (bit-exact)."""
    g = torch.Generator().manual_seed(5)
    x = torch.rand(5, 4, 24, 40, generator=g).to(dtype)
    mask = torch.tensor([True, False, True, True, False])
    colors = torch.rand(5, 3, generator=g) * 0.6 + 0.3
    aug = P.RandomBackgroundBlend(prob=0.5, color_range=(0.3, 0.9))
    out, flag = aug(x.cuda(), mask=mask, colors=colors)
    assert flag.cpu().tolist() == mask.tolist()
    for i in range(5):
        want = O.background_blend(x[i], colors[i]) if mask[i] else x[i]
        assert torch.equal(out[i].cpu(), want), i
    # random path: the fraction of augmented samples follows prob, colours stay inside the range, alpha becomes 1
    big = torch.rand(256, 4, 4, 4, generator=g)
    big[:, 3] = 0.0  # fully transparent: the output RGB is exactly the drawn colour
    out, flag = P.RandomBackgroundBlend(prob=0.25, color_range=(0.3, 0.9))(big.cuda(), generator=torch.Generator("cuda").manual_seed(7))
    frac = float(flag.float().mean())
    assert 0.12 < frac < 0.40
    sel = out[flag]
    assert float(sel[:, 3].min()) == 1.0 and 0.3 <= float(sel[:, :3].min()) and float(sel[:, :3].max()) <= 0.9
    assert torch.equal(out[~flag].cpu(), big[~flag.cpu()])
    with pytest.raises(ValueError):
        P.RandomBackgroundBlend(color_range=(0.9, 0.2))


def test_build_training_batch(P):
    g = torch.Generator().manual_seed(9)
    comp, full, bgd = (torch.rand(2, 4, 8, 8, generator=g) for _ in range(3))
    got = P.build_training_batch({"component": comp, "composite": full}, "cuda")
    assert torch.equal(got.cpu(), O.build_training_batch({"component": comp, "composite": full}))
    got = P.build_training_batch({"composite": full, "background": bgd}, "cuda", background_sample_prob=1.0)
    assert torch.equal(got.cpu(), O.build_training_batch({"composite": full, "background": bgd}, torch.tensor([True, True])))
    got = P.build_training_batch({"composite": full, "background": bgd[0]}, "cuda", background_sample_prob=0.0)
    assert got.shape[0] == 2
    with pytest.raises(ValueError):
        P.build_training_batch({"component": comp}, "cuda")
    with pytest.raises(ValueError):
        P.build_training_batch({"composite": full, "background": bgd[:, :3]}, "cuda", background_sample_prob=1.0)
