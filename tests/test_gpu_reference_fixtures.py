"""GPU parity of the CUDA path against fixtures the REFERENCE'S OWN functions produced.

``tests/golden/ref_functions.safetensors`` / ``ref_forward_<arch>.safetensors`` / ``ref_manifest.json`` come from
scripts/make_reference_fixtures.py (the unmodified reference files executed through tests/refshim.py); the reference
itself cannot travel to the GPU box.  fp32 inputs: kernels accumulate in fp32/fp64, so loss / PSNR terms are held to
1e-5 relative; elementwise outputs (triplet, blend, composite) to exact equality or 1 ulp where the reference's own
operation order is kept.  bf16 forward: north-star tolerances (2e-2, 0.05 dB)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def R(lib_built):
    import ragb_vae_b200 as r

    assert torch.cuda.is_available()
    return r


@pytest.fixture(scope="module")
def fx():
    from safetensors.torch import load_file

    return {k: v.cuda() for k, v in load_file(os.path.join(GOLD, "ref_functions.safetensors")).items()}


@pytest.fixture(scope="module")
def man():
    with open(os.path.join(GOLD, "ref_manifest.json")) as f:
        return json.load(f)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_recon_loss_kernel_matches_reference_alpha_vae_loss(R, fx, man):
    pred, target = fx["loss_pred"], fx["loss_target"]
    for rm in (False, True):
        for naive in (False, True):
            ref = man["recon_loss"][f"reduce_mean={rm},naive={naive}"]
            got = float(R.AlphaVaeLoss(reduce_mean=rm, use_naive_mse=naive).reconstruction_loss(pred, target))
            assert got == pytest.approx(ref, rel=1e-5), (rm, naive)
    c = man["recon_loss_custom_eb"]
    got = float(R.AlphaVaeLoss(reduce_mean=True, custom_eb=c["eb"], custom_eb2=c["eb2"]).reconstruction_loss(pred, target))
    assert got == pytest.approx(c["value"], rel=1e-5)
    with pytest.raises(ValueError) as e:
        R.AlphaVaeLoss(custom_eb=(1.0, 2.0))
    assert [type(e.value).__name__, str(e.value)] == man["errors"]["loss_bad_eb"]


def test_posterior_kl_kernels_match_reference_kl_loss(R, fx, man):
    p = R.DiagonalGaussianDistribution(fx["kl_moments"])
    q = R.DiagonalGaussianDistribution(fx["kl_moments_other"])
    for rm in (False, True):
        mod = R.AlphaVaeLoss(reduce_mean=rm)
        assert float(mod.kl_loss(p)) == pytest.approx(man["kl_loss"][f"reduce_mean={rm}"], rel=1e-5)
        assert float(mod.kl_loss(p, q)) == pytest.approx(man["kl_loss"][f"reduce_mean={rm},other"], rel=1e-5)


def test_composite_mirror_and_fused_psnr_match_reference(R, fx, man):
    rgba = fx["comp_rgba"]
    cases = {"comp_white": 1.0, "comp_black": 0.0, "comp_grey": 0.3, "comp_triple": (0.2, 0.5, 0.9), "comp_tensor3": fx["comp_bg3"],
             "comp_tensor4": fx["comp_bg4"], "comp_tensor1": fx["comp_bg1"]}
    for key, bg in cases.items():
        assert torch.equal(R.composite_over_background(rgba, bg), fx[key]), key
    assert torch.equal(R.composite_over_background(rgba[:, :3], (0.2, 0.5, 0.9)), fx["comp_rgb_only"])
    for key, bad in (("comp_two_values", (1.0, 0.0)), ("comp_bad_rank", torch.zeros(24, 32)), ("comp_bad_size", torch.zeros(3, 4, 4))):
        with pytest.raises(ValueError) as e:
            R.composite_over_background(rgba, bad)
        assert str(e.value) == man["errors"][key][1], key
    # the fused composite + PSNR + alpha-MAE kernel against compute_psnr(composite(recon), composite(gt))
    m = R.validation_metrics(fx["val_recon"], rgba, backgrounds=("white", "black", (0.2, 0.5, 0.9)))
    assert torch.allclose(m["psnr_white"], fx["val_psnr_white"], atol=1e-4)
    assert torch.allclose(m["psnr_black"], fx["val_psnr_black"], atol=1e-4)
    assert torch.allclose(m["psnr_(0.2, 0.5, 0.9)"], fx["val_psnr_triple"], atol=1e-4)
    assert torch.allclose(m["alpha_mae"], fx["val_alpha_mae"], rtol=1e-5, atol=1e-7)
    got = R.compute_psnr(fx["psnr_pred"], fx["psnr_target"])
    assert torch.allclose(got, fx["psnr_out"], atol=1e-4) and float(got[2]) == pytest.approx(80.0, abs=1e-4)
    from ragb_vae_b200.validation import resolve_background_spec

    assert resolve_background_spec("white") == man["background_spec"]["white"] and resolve_background_spec("BLACK") == 0.0
    with pytest.raises(ValueError) as e:
        resolve_background_spec("green")
    assert str(e.value) == man["errors"]["bad_background_spec"][1]


def test_triplet_blend_batch_kernels_match_reference(R, fx, man):
    from ragb_vae_b200 import plumbing as P

    out = P.build_detail_augmented_triplet(fx["triplet_in"])
    assert out.shape == fx["triplet_out"].shape and float((out - fx["triplet_out"]).abs().max()) <= 1.2e-7
    with pytest.raises(ValueError) as e:
        P.build_detail_augmented_triplet(fx["triplet_in"][:, :3])
    assert str(e.value) == man["errors"]["triplet_rgb"][1]
    parts = P.split_triplet_distribution(R.DiagonalGaussianDistribution(fx["split_in"]))
    for i, p in enumerate(parts):
        assert torch.equal(p.parameters, fx[f"split_out{i}"])
    with pytest.raises(ValueError) as e:
        P.split_triplet_distribution(R.DiagonalGaussianDistribution(fx["split_in"][:2]))
    assert str(e.value) == man["errors"]["split_not_triplet"][1]
    blend = P.RandomBackgroundBlend(prob=1.0)
    got, flagged = blend(fx["blend_in"][None], colors=fx["blend_color"][None], mask=torch.ones(1, dtype=torch.bool, device="cuda"))
    assert bool(flagged[0]) and float((got[0] - fx["blend_out"]).abs().max()) <= 1.2e-7
    with pytest.raises(ValueError) as e:
        P.RandomBackgroundBlend(color_range=(0.9, 0.2))
    assert str(e.value) == man["errors"]["blend_bad_range"][1]
    batch = {k: fx[f"batch_{k}"] for k in ("component", "composite", "background")}
    assert torch.equal(P.build_training_batch(batch, torch.device("cuda"), background_mask=fx["batch_mask"].bool()), fx["batch_out_bg"])
    assert torch.equal(P.build_training_batch(batch, torch.device("cuda")), fx["batch_out_plain"])
    assert torch.equal(P.build_training_batch({"composite": batch["composite"]}, torch.device("cuda")), fx["batch_out_composite_only"])
    with pytest.raises(ValueError) as e:
        P.build_training_batch({"component": batch["component"]}, torch.device("cuda"))
    assert str(e.value) == man["errors"]["batch_no_composite"][1]


@pytest.mark.parametrize("arch", ["qwen", "flux"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_loss_and_eval_match_reference_rgba_vae(R, man, oracle_model, arch, dtype):
    """R.RgbaVAE.forward / .loss / evaluate_rgba_vae on the GPU against what the reference's RgbaVAE.forward / .loss /
    evaluate_rgba_vae produced over the oracle VAE with the same weights and the recorded eps."""
    from safetensors.torch import load_file

    g = {k: v.cuda() for k, v in load_file(os.path.join(GOLD, f"ref_forward_{arch}.safetensors")).items()}
    entry = [e for e in man["forward"] if e["arch"] == arch][0]
    vae = R.RgbaAutoencoder(arch)
    vae.load_state_dict(oracle_model(arch).state_dict())
    vae = vae.to("cuda", dtype)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    model = R.RgbaVAE(vae)
    recon, post = model(g["x"].to(dtype), noise=g["noise"])
    assert rel(recon, g["recon"]) < tol and rel(post.parameters, g["moments"]) < tol
    recon3, post3 = model(g["x"][:1, :3].to(dtype), noise=g["noise3"])
    assert rel(recon3, g["recon3"]) < tol and rel(post3.parameters, g["moments3"]) < tol
    # RgbaVAE.loss: the weighted terms on the REFERENCE's (recon, target, posterior) -- isolates the loss kernels
    ref_post = R.DiagonalGaussianDistribution(g["moments"])
    for name, kw in man["loss_configs"].items():
        got = float(R.RgbaVAE(vae, **kw).loss(g["recon"], g["x"], ref_post))
        assert got == pytest.approx(entry["loss"][name], rel=2e-5), name
    got = float(model.loss(g["recon3"], g["x"][:1, :3], R.DiagonalGaussianDistribution(g["moments3"])))
    assert got == pytest.approx(entry["loss_rgb_target"], rel=2e-5)
    # ... and end to end on our own reconstruction
    for name in ("flux_vae_yaml", "all_terms"):
        got = float(R.RgbaVAE(vae, **man["loss_configs"][name]).loss(recon, g["x"].to(dtype), post))
        assert got == pytest.approx(entry["loss"][name], rel=5e-4 if dtype == torch.float32 else 5e-2), name
    # the validation loop: means the reference printed (2 decimals for PSNR, 4 for alpha MAE)
    res = R.evaluate_rgba_vae(model, [g["eval_batch0"].to(dtype), g["eval_batch1"].to(dtype)],
                              backgrounds=("white", "black", (0.2, 0.5, 0.9)), noises=[g["eval_noise0"], g["eval_noise1"]])
    printed = [float(line.split(": ")[1].split(" ")[0]) for line in entry["eval_lines"]]
    keys = ["psnr_white", "psnr_black", "psnr_(0.2, 0.5, 0.9)"]
    for k, p in zip(keys, printed[:3]):
        assert abs(res[k] - p) < 0.005 + (1e-3 if dtype == torch.float32 else 0.05), (k, res[k], p)
    assert abs(res["alpha_mae"] - printed[3]) < 5e-5 + (1e-5 if dtype == torch.float32 else 5e-3)


def test_adapt_vae_to_rgba_mirror_matches_reference(R, fx):
    from types import SimpleNamespace

    for tag, conv in (("2d", torch.nn.Conv2d), ("3d", torch.nn.Conv3d)):
        holder = SimpleNamespace(encoder=SimpleNamespace(conv_in=conv(3, 8, 3).cuda()), decoder=SimpleNamespace(conv_out=conv(8, 3, 3).cuda()),
                                 config=SimpleNamespace(in_channels=3, out_channels=3))
        with torch.no_grad():
            holder.encoder.conv_in.weight.copy_(fx[f"adapt{tag}_in_w"])
            holder.encoder.conv_in.bias.copy_(fx[f"adapt{tag}_in_b"])
            holder.decoder.conv_out.weight.copy_(fx[f"adapt{tag}_out_w"])
            holder.decoder.conv_out.bias.copy_(fx[f"adapt{tag}_out_b"])
        R.adapt_vae_to_rgba(holder, alpha_bias_init=0.7)
        assert torch.equal(holder.encoder.conv_in.weight, fx[f"adapt{tag}_in_w4"])
        assert torch.equal(holder.decoder.conv_out.weight, fx[f"adapt{tag}_out_w4"])
        assert torch.equal(holder.decoder.conv_out.bias, fx[f"adapt{tag}_out_b4"])
        assert holder.config.in_channels == 4 and holder.config.out_channels == 4
