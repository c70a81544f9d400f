"""CPU tests that pin the oracle to the REFERENCE'S OWN functions.

``tests/golden/ref_*.safetensors`` / ``ref_manifest.json`` were written by scripts/make_reference_fixtures.py from the
unmodified reference files (src/models/losses.py, src/models/rgba_vae.py, src/training/rgba_vae_stage.py) executed through
the ``sys.modules`` shim of tests/refshim.py.  Here (1) every oracle restatement of those functions is compared with the
fixtures (runs everywhere), and (2) when ``/root/reference`` is present the fixtures are re-derived live from the reference
and must be unchanged (so a stale fixture cannot hide a drift).  The GPU twins are in tests/test_gpu_reference_fixtures.py.
"""
import json
import os

import pytest
import torch

import refshim
from oracle import vae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def fx():
    from safetensors.torch import load_file

    return load_file(os.path.join(GOLD, "ref_functions.safetensors"))


@pytest.fixture(scope="module")
def man():
    with open(os.path.join(GOLD, "ref_manifest.json")) as f:
        return json.load(f)


def close(a, b, tol=1e-6):
    return float((a.double() - b.double()).abs().max()) <= tol * max(1.0, float(b.double().abs().max()))


# ---- (1) oracle == reference fixtures --------------------------------------------------------------------------
def test_oracle_reconstruction_loss_matches_reference(fx, man):
    pred, target = fx["loss_pred"], fx["loss_target"]
    for rm in (False, True):
        for naive in (False, True):
            ref = man["recon_loss"][f"reduce_mean={rm},naive={naive}"]
            got = float(O.reconstruction_loss(pred, target, reduce_mean=rm, use_naive_mse=naive))
            assert got == pytest.approx(ref, rel=1e-6), (rm, naive)
    c = man["recon_loss_custom_eb"]
    got = float(O.reconstruction_loss(pred, target, reduce_mean=True, eb=tuple(c["eb"]), eb2=tuple(c["eb2"])))
    assert got == pytest.approx(c["value"], rel=1e-6)


def test_oracle_kl_loss_matches_reference(fx, man):
    p, q = O.DiagonalGaussianDistribution(fx["kl_moments"]), O.DiagonalGaussianDistribution(fx["kl_moments_other"])
    for rm in (False, True):
        assert float(O.kl_loss(p, None, rm)) == pytest.approx(man["kl_loss"][f"reduce_mean={rm}"], rel=1e-6)
        assert float(O.kl_loss(p, q, rm)) == pytest.approx(man["kl_loss"][f"reduce_mean={rm},other"], rel=1e-6)


def test_oracle_composite_matches_reference(fx, man):
    rgba = fx["comp_rgba"]
    cases = {"comp_white": 1.0, "comp_black": 0.0, "comp_grey": 0.3, "comp_triple": (0.2, 0.5, 0.9), "comp_tensor3": fx["comp_bg3"],
             "comp_tensor4": fx["comp_bg4"], "comp_tensor1": fx["comp_bg1"]}
    for key, bg in cases.items():
        assert torch.equal(O.composite_over_background(rgba, bg), fx[key]), key
    assert torch.equal(O.composite_over_background(rgba[:, :3], (0.2, 0.5, 0.9)), fx["comp_rgb_only"])
    for key, bad in (("comp_two_values", (1.0, 0.0)), ("comp_bad_rank", torch.zeros(24, 32)), ("comp_bad_size", torch.zeros(3, 4, 4))):
        kind, msg = man["errors"][key]
        with pytest.raises(ValueError) as e:
            O.composite_over_background(rgba, bad)
        assert kind == "ValueError" and str(e.value) == msg


def test_oracle_psnr_and_validation_body_match_reference(fx):
    assert close(O.compute_psnr(fx["psnr_pred"], fx["psnr_target"]), fx["psnr_out"])
    assert float(fx["psnr_out"][2]) == pytest.approx(80.0, abs=1e-4)  # identical pair: mse clamp 1e-8
    m = O.validation_metrics(fx["val_recon"], fx["comp_rgba"], backgrounds=(1.0, 0.0, (0.2, 0.5, 0.9)))
    assert close(m[1.0], fx["val_psnr_white"]) and close(m[0.0], fx["val_psnr_black"])
    assert close(m[(0.2, 0.5, 0.9)], fx["val_psnr_triple"]) and close(m["alpha_mae"], fx["val_alpha_mae"])


def test_oracle_triplet_split_blend_batch_match_reference(fx, man):
    assert torch.equal(O.build_detail_augmented_triplet(fx["triplet_in"]), fx["triplet_out"])
    with pytest.raises(ValueError) as e:
        O.build_detail_augmented_triplet(fx["triplet_in"][:, :3])
    assert str(e.value) == man["errors"]["triplet_rgb"][1]
    parts = O.split_triplet_distribution(O.DiagonalGaussianDistribution(fx["split_in"]))
    for i, p in enumerate(parts):
        assert torch.equal(p.parameters, fx[f"split_out{i}"])
    assert torch.equal(O.background_blend(fx["blend_in"], fx["blend_color"]), fx["blend_out"])
    batch = {k: fx[f"batch_{k}"] for k in ("component", "composite", "background")}
    assert torch.equal(O.build_training_batch(batch, fx["batch_mask"].bool()), fx["batch_out_bg"])
    assert torch.equal(O.build_training_batch(batch), fx["batch_out_plain"])
    assert torch.equal(O.build_training_batch({"composite": batch["composite"]}), fx["batch_out_composite_only"])
    with pytest.raises(ValueError) as e:
        O.build_training_batch({"component": batch["component"]})
    assert str(e.value) == man["errors"]["batch_no_composite"][1]


@pytest.mark.parametrize("tag,conv", [("2d", torch.nn.Conv2d), ("3d", torch.nn.Conv3d)])
def test_oracle_adapt_vae_to_rgba_matches_reference(fx, tag, conv):
    from types import SimpleNamespace

    holder = SimpleNamespace(encoder=SimpleNamespace(conv_in=conv(3, 8, 3)), decoder=SimpleNamespace(conv_out=conv(8, 3, 3)),
                             config=SimpleNamespace(in_channels=3, out_channels=3))
    with torch.no_grad():
        holder.encoder.conv_in.weight.copy_(fx[f"adapt{tag}_in_w"])
        holder.encoder.conv_in.bias.copy_(fx[f"adapt{tag}_in_b"])
        holder.decoder.conv_out.weight.copy_(fx[f"adapt{tag}_out_w"])
        holder.decoder.conv_out.bias.copy_(fx[f"adapt{tag}_out_b"])
    O.adapt_vae_to_rgba(holder, alpha_bias_init=0.7)
    assert torch.equal(holder.encoder.conv_in.weight, fx[f"adapt{tag}_in_w4"])
    assert torch.equal(holder.encoder.conv_in.bias, fx[f"adapt{tag}_in_b4"])
    assert torch.equal(holder.decoder.conv_out.weight, fx[f"adapt{tag}_out_w4"])
    assert torch.equal(holder.decoder.conv_out.bias, fx[f"adapt{tag}_out_b4"])
    assert holder.config.in_channels == 4 and holder.config.out_channels == 4


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_oracle_forward_matches_reference_rgba_vae_forward(arch, man, oracle_model):
    """O.rgba_vae_forward == the reference's RgbaVAE.forward (rgba_vae.py:274-281) run over the same VAE and eps."""
    from safetensors.torch import load_file

    g = load_file(os.path.join(GOLD, f"ref_forward_{arch}.safetensors"))
    vae = oracle_model(arch)
    entry = [e for e in man["forward"] if e["arch"] == arch][0]
    assert float(sum(p.double().abs().sum() for p in vae.parameters())) == pytest.approx(entry["weight_checksum"], rel=1e-9)
    recon, post, _ = O.rgba_vae_forward(vae, g["x"], g["noise"])
    assert close(recon, g["recon"], 1e-5) and close(post.parameters, g["moments"], 1e-5)
    recon3, post3, _ = O.rgba_vae_forward(vae, g["x"][:1, :3], g["noise3"])
    assert close(recon3, g["recon3"], 1e-5) and close(post3.parameters, g["moments3"], 1e-5)
    # the validation loop's printed means (evaluate_rgba_vae, rgba_vae_stage.py:766-774)
    rows = []
    for b, n in ((g["eval_batch0"], g["eval_noise0"]), (g["eval_batch1"], g["eval_noise1"])):
        r, _, _ = O.rgba_vae_forward(vae, b, n)
        m = O.validation_metrics(r, b, backgrounds=(1.0, 0.0, (0.2, 0.5, 0.9)))
        rows.append(torch.stack([m[1.0], m[0.0], m[(0.2, 0.5, 0.9)], m["alpha_mae"]], 1))
    mean = torch.cat(rows).mean(0)
    printed = [float(line.split(": ")[1].split(" ")[0]) for line in entry["eval_lines"]]
    assert [round(float(v), 2) for v in mean[:3]] == pytest.approx(printed[:3], abs=0.011)
    assert float(mean[3]) == pytest.approx(printed[3], abs=6e-5)


# ---- (2) fixtures == the reference, live ------------------------------------------------------------------------
@pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference is not present on this box")
def test_fixtures_are_what_the_reference_computes_today(fx, man, tmp_path, monkeypatch):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_reference_fixtures", os.path.join(ROOT, "scripts", "make_reference_fixtures.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    ref = refshim.load_reference(gen.RecordingVAE, O.DiagonalGaussianDistribution)
    t, m = {}, {}
    gen.functions(ref, t, m)
    assert set(t) == set(fx)
    for k in t:
        assert torch.equal(t[k], fx[k]), k
    for k in ("recon_loss", "recon_loss_custom_eb", "kl_loss", "errors", "background_spec"):
        assert json.loads(json.dumps(m[k])) == man[k], k
