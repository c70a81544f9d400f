#!/usr/bin/env python
"""Generate tests/golden/ref_*.safetensors + ref_manifest.json from the REFERENCE'S OWN functions.

Run once in the build container (``/root/reference`` present).  The unmodified reference files are executed
where they lie through the ``sys.modules`` shim of tests/refshim.py; the only substituted names are
``diffusers.AutoencoderKL`` / ``DiagonalGaussianDistribution`` (diffusers is not installable here), which are
bound to the oracle's restatement of the diffusers classes.  So:

* **reference-pinned** (pure reference code, no oracle involved in the arithmetic):
  ``AlphaVaeLoss.reconstruction_loss`` / ``_reduce`` (src/models/losses.py:67-83,117-123),
  ``composite_over_background`` / ``_normalize_background`` / ``_ensure_alpha`` (src/models/rgba_vae.py:25-92),
  ``adapt_vae_to_rgba`` (:95-123), ``compute_psnr`` (src/training/rgba_vae_stage.py:712-715),
  ``build_detail_augmented_triplet`` (:606-625), ``RandomBackgroundBlend._blend_tensor`` (:118-129),
  ``build_training_batch`` (:575-603), ``resolve_background_spec`` (:787-795), and the weighted terms of
  ``RgbaVAE.loss`` (rgba_vae.py:283-316) given (recon, target, posterior);
* **reference code over the oracle's diffusers restatement** (the orchestration is the reference's, the conv /
  norm / attention arithmetic and the posterior are the oracle's -- that half stays unpinned):
  ``RgbaVAE.forward`` (rgba_vae.py:274-281), ``AlphaVaeLoss.kl_loss`` (losses.py:109-115),
  ``split_triplet_distribution`` (rgba_vae_stage.py:690-700), the ``evaluate_rgba_vae`` loop (:718-784).

The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import io
import json
import os
import sys
from contextlib import redirect_stdout
from types import SimpleNamespace

import torch
import torch.nn as nn
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refshim  # noqa: E402
from oracle import vae_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class RecordingGaussian(O.DiagonalGaussianDistribution):
    """The oracle posterior, remembering every eps ``sample()`` drew (the reference never passes noise)."""
    drawn = []

    def sample(self, generator=None, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype)
            RecordingGaussian.drawn.append(noise.clone())
        return super().sample(generator=generator, noise=noise)


class RecordingVAE(O.OracleVAE):
    def encode(self, x):
        return SimpleNamespace(latent_dist=RecordingGaussian(self.encode_moments(x)))


def rand(shape, seed, lo=0.0, hi=1.0):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed)) * (hi - lo) + lo


def error_of(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        return [type(e).__name__, str(e)]
    return None


def functions(ref, t, man):
    L, RV, ST = ref.losses, ref.rgba_vae, ref.stage
    # --- AlphaVaeLoss.reconstruction_loss / _reduce ------------------------------------------------
    pred, target = rand((2, 4, 40, 56), 11, -1, 1), rand((2, 4, 40, 56), 12, -1, 1)
    t["loss_pred"], t["loss_target"] = pred, target
    man["recon_loss"] = {}
    for rm in (False, True):
        for naive in (False, True):
            v = L.AlphaVaeLoss(reduce_mean=rm, use_naive_mse=naive).reconstruction_loss(pred, target)
            man["recon_loss"][f"reduce_mean={rm},naive={naive}"] = float(v)
    eb, eb2 = (0.1, -0.2, 0.3), (0.5, 0.25, 0.125)
    man["recon_loss_custom_eb"] = {"eb": eb, "eb2": eb2, "value": float(
        L.AlphaVaeLoss(reduce_mean=True, custom_eb=eb, custom_eb2=eb2).reconstruction_loss(pred, target))}
    man["errors"] = {"loss_bad_eb": error_of(lambda: L.AlphaVaeLoss(custom_eb=(1.0, 2.0)))}
    # --- kl_loss (reference _reduce over the posterior class bound by the shim) --------------------
    m1 = torch.randn(2, 32, 8, 8, generator=torch.Generator().manual_seed(13))
    m2 = torch.randn(2, 32, 8, 8, generator=torch.Generator().manual_seed(14))
    m1[0, 16, 0, 0], m1[0, 17, 0, 0] = 45.0, -45.0  # exercise the logvar clamp(-30, 20)
    t["kl_moments"], t["kl_moments_other"] = m1, m2
    man["kl_loss"] = {}
    for rm in (False, True):
        mod = L.AlphaVaeLoss(reduce_mean=rm)
        p, q = O.DiagonalGaussianDistribution(m1), O.DiagonalGaussianDistribution(m2)
        man["kl_loss"][f"reduce_mean={rm}"] = float(mod.kl_loss(p))
        man["kl_loss"][f"reduce_mean={rm},other"] = float(mod.kl_loss(p, q))
    # --- composite_over_background ------------------------------------------------------------------
    rgba = rand((2, 4, 24, 32), 15)
    rgba[:, 3, :8] = 0.0
    rgba[:, 3, 16:] = 1.0
    t["comp_rgba"] = rgba
    bg3, bg4, bg1 = rand((3, 24, 32), 16), rand((2, 3, 24, 32), 17), rand((1, 1, 24, 32), 18)
    t["comp_bg3"], t["comp_bg4"], t["comp_bg1"] = bg3, bg4, bg1
    t["comp_white"] = RV.composite_over_white(rgba)
    t["comp_black"] = RV.composite_over_black(rgba)
    t["comp_grey"] = RV.composite_over_background(rgba, 0.3)
    t["comp_triple"] = RV.composite_over_background(rgba, (0.2, 0.5, 0.9))
    t["comp_tensor3"] = RV.composite_over_background(rgba, bg3)
    t["comp_tensor4"] = RV.composite_over_background(rgba, bg4)
    t["comp_tensor1"] = RV.composite_over_background(rgba, bg1)
    t["comp_rgb_only"] = RV.composite_over_background(rgba[:, :3], (0.2, 0.5, 0.9))  # _ensure_alpha: alpha = 1
    man["errors"]["comp_two_values"] = error_of(lambda: RV.composite_over_background(rgba, (1.0, 0.0)))
    man["errors"]["comp_bad_rank"] = error_of(lambda: RV.composite_over_background(rgba, torch.zeros(24, 32)))
    man["errors"]["comp_bad_size"] = error_of(lambda: RV.composite_over_background(rgba, torch.zeros(3, 4, 4)))
    # --- compute_psnr / resolve_background_spec -----------------------------------------------------
    pp, pt = rand((3, 3, 24, 32), 19), rand((3, 3, 24, 32), 20)
    pp[2] = pt[2]  # identical pair: mse clamps at 1e-8 -> 80 dB
    t["psnr_pred"], t["psnr_target"], t["psnr_out"] = pp, pt, ST.compute_psnr(pp, pt)
    man["background_spec"] = {"white": ST.resolve_background_spec("white"), "BLACK": ST.resolve_background_spec("BLACK"),
                              "triple": list(ST.resolve_background_spec((0.1, 0.2, 0.3)))}
    man["errors"]["bad_background_spec"] = error_of(lambda: ST.resolve_background_spec("green"))
    # --- validation body: composite both, PSNR, alpha MAE (rgba_vae_stage.py:742-753) ---------------
    recon = (rgba + 0.05 * torch.randn(rgba.shape, generator=torch.Generator().manual_seed(21))).clamp(0, 1)
    t["val_recon"] = recon
    for name, bg in (("white", 1.0), ("black", 0.0), ("triple", (0.2, 0.5, 0.9))):
        t[f"val_psnr_{name}"] = ST.compute_psnr(RV.composite_over_background(recon, bg), RV.composite_over_background(rgba, bg))
    t["val_alpha_mae"] = torch.mean(torch.abs(recon[:, 3:] - rgba[:, 3:]), dim=(1, 2, 3))
    # --- triplet / split ----------------------------------------------------------------------------
    tgt = rand((2, 4, 24, 32), 22, -1, 1)
    t["triplet_in"], t["triplet_out"] = tgt, ST.build_detail_augmented_triplet(tgt)
    man["errors"]["triplet_rgb"] = error_of(lambda: ST.build_detail_augmented_triplet(tgt[:, :3]))
    m6 = torch.randn(6, 32, 4, 4, generator=torch.Generator().manual_seed(23))
    parts = ST.split_triplet_distribution(O.DiagonalGaussianDistribution(m6))
    t["split_in"] = m6
    for i, p in enumerate(parts):
        t[f"split_out{i}"] = p.parameters.clone()
    man["errors"]["split_not_triplet"] = error_of(
        lambda: ST.split_triplet_distribution(O.DiagonalGaussianDistribution(m6[:2])))
    # --- RandomBackgroundBlend._blend_tensor (colour drawn from the global RNG) ---------------------
    comp = rand((4, 24, 32), 24)
    torch.manual_seed(7)
    t["blend_in"], t["blend_out"] = comp, ST.RandomBackgroundBlend(prob=1.0)._blend_tensor(comp)
    torch.manual_seed(7)
    t["blend_color"] = torch.empty((3, 1, 1)).uniform_(0.2, 0.9).reshape(3)
    man["errors"]["blend_bad_range"] = error_of(lambda: ST.RandomBackgroundBlend(color_range=(0.9, 0.2)))
    # --- build_training_batch -----------------------------------------------------------------------
    batch = {"component": rand((2, 4, 16, 16), 25), "composite": rand((2, 4, 16, 16), 26), "background": rand((3, 4, 16, 16), 27)}
    for k, v in batch.items():
        t[f"batch_{k}"] = v
    torch.manual_seed(11)
    t["batch_out_bg"] = ST.build_training_batch(batch, torch.device("cpu"), background_sample_prob=0.6)
    torch.manual_seed(11)
    t["batch_mask"] = (torch.rand(3) < 0.6).to(torch.uint8)
    t["batch_out_plain"] = ST.build_training_batch(batch, torch.device("cpu"))
    t["batch_out_composite_only"] = ST.build_training_batch({"composite": batch["composite"]}, torch.device("cpu"))
    man["errors"]["batch_no_composite"] = error_of(lambda: ST.build_training_batch({"component": batch["component"]}, torch.device("cpu")))
    man["errors"]["batch_rgb_background"] = error_of(lambda: ST.build_training_batch(
        {"composite": batch["composite"], "background": batch["background"][:, :3]}, torch.device("cpu"), background_sample_prob=1.0))
    # --- adapt_vae_to_rgba on 2-D and 3-D (rank-agnostic) convs -------------------------------------
    for tag, conv in (("2d", nn.Conv2d), ("3d", nn.Conv3d)):
        torch.manual_seed(31)
        holder = SimpleNamespace(encoder=SimpleNamespace(conv_in=conv(3, 8, 3)), decoder=SimpleNamespace(conv_out=conv(8, 3, 3)),
                                 config=SimpleNamespace(in_channels=3, out_channels=3))
        t[f"adapt{tag}_in_w"], t[f"adapt{tag}_in_b"] = holder.encoder.conv_in.weight.data.clone(), holder.encoder.conv_in.bias.data.clone()
        t[f"adapt{tag}_out_w"], t[f"adapt{tag}_out_b"] = holder.decoder.conv_out.weight.data.clone(), holder.decoder.conv_out.bias.data.clone()
        RV.adapt_vae_to_rgba(holder, alpha_bias_init=0.7)
        t[f"adapt{tag}_in_w4"], t[f"adapt{tag}_in_b4"] = holder.encoder.conv_in.weight.data.clone(), holder.encoder.conv_in.bias.data.clone()
        t[f"adapt{tag}_out_w4"], t[f"adapt{tag}_out_b4"] = holder.decoder.conv_out.weight.data.clone(), holder.decoder.conv_out.bias.data.clone()
        assert holder.config.in_channels == 4 and holder.config.out_channels == 4
        assert holder.encoder.conv_in.in_channels == 4 and holder.decoder.conv_out.out_channels == 4


LOSS_CONFIGS = {
    "default": {},  # alpha_loss_weight = 1 (rgba_vae.py:199)
    "flux_vae_yaml": dict(white_bg_weight=0.5, black_bg_weight=0.5, loss_reduce_mean=True),
    "all_terms": dict(beta=0.5, alpha_loss_weight=0.7, alpha_l1_weight=0.3, rgb_loss_weight=1.5, white_bg_weight=0.25,
                      black_bg_weight=0.125, loss_reduce_mean=True),
    "naive_mse": dict(use_naive_mse=True, alpha_loss_weight=0.0, loss_reduce_mean=True),
    "no_rgb_term": dict(rgb_loss_weight=0.0, alpha_l1_weight=1.0),
    "custom_eb": dict(custom_eb=(0.1, -0.2, 0.3), custom_eb2=(0.5, 0.25, 0.125), alpha_loss_weight=0.0),
}


def forward(ref, arch, man):
    """The reference RgbaVAE (forward / loss / evaluate loop) over the oracle VAE, 64x64."""
    torch.manual_seed(0)
    vae = RecordingVAE(arch, 4, 4).eval().requires_grad_(False)
    check = O.build_oracle(arch, seed=0)
    assert all(torch.equal(a, b) for a, b in zip(vae.state_dict().values(), check.state_dict().values()))
    t = {}
    model = ref.rgba_vae.RgbaVAE(vae=vae)
    x = O.synthetic_rgba(2, 64, 64, seed=41, structured=True)
    RecordingGaussian.drawn.clear()
    torch.manual_seed(42)
    with torch.no_grad():
        recon, post = model(x)
        recon3, post3 = model(x[:1, :3])  # 3-channel input: _ensure_alpha
    t["x"], t["noise"], t["noise3"] = x, RecordingGaussian.drawn[0], RecordingGaussian.drawn[1]
    t["recon"], t["moments"], t["recon3"], t["moments3"] = recon, post.parameters, recon3, post3.parameters
    entry = {"arch": arch, "loss": {}, "weight_checksum": float(sum(p.double().abs().sum() for p in vae.parameters()))}
    for name, kw in LOSS_CONFIGS.items():
        m = ref.rgba_vae.RgbaVAE(vae=vae, **kw)
        entry["loss"][name] = float(m.loss(recon, x, post))
    entry["loss_rgb_target"] = float(model.loss(recon3, x[:1, :3], post3))
    # use_naive_mse with the per-sample-sum reduction: the reference calls .view on a channel slice and raises
    entry["loss_naive_mse_sum_error"] = error_of(lambda: ref.rgba_vae.RgbaVAE(vae=vae, use_naive_mse=True).loss(recon, x, post))
    # evaluate_rgba_vae with a stand-in accelerator (single process: gather = identity)
    lines = []
    acc = SimpleNamespace(device=torch.device("cpu"), gather=lambda v: v, print=lambda s: lines.append(s), is_main_process=False)
    batches = [{"composite": O.synthetic_rgba(2, 64, 64, seed=43, structured=True)}, {"composite": O.synthetic_rgba(1, 64, 64, seed=44)}]
    RecordingGaussian.drawn.clear()
    torch.manual_seed(45)
    with redirect_stdout(io.StringIO()):
        ref.stage.evaluate_rgba_vae(acc, model, batches, epoch=3, eval_cfg={"val_background_colors": ["white", "black", (0.2, 0.5, 0.9)]})
    t["eval_batch0"], t["eval_batch1"] = batches[0]["composite"], batches[1]["composite"]
    t["eval_noise0"], t["eval_noise1"] = RecordingGaussian.drawn[0], RecordingGaussian.drawn[1]
    entry["eval_lines"] = lines
    man["forward"].append(entry)
    save_file({k: v.contiguous() for k, v in t.items()}, os.path.join(GOLD, f"ref_forward_{arch}.safetensors"))
    print(arch, json.dumps(entry, indent=1))


def main():
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    ref = refshim.load_reference(RecordingVAE, O.DiagonalGaussianDistribution)
    man = {"torch": torch.__version__, "made_by": "scripts/make_reference_fixtures.py",
           "reference_files": ["src/models/losses.py", "src/models/rgba_vae.py", "src/training/rgba_vae_stage.py"],
           "loss_configs": {k: {a: (list(b) if isinstance(b, tuple) else b) for a, b in v.items()} for k, v in LOSS_CONFIGS.items()},
           "forward": []}
    t = {}
    functions(ref, t, man)
    save_file({k: v.contiguous() for k, v in t.items()}, os.path.join(GOLD, "ref_functions.safetensors"))
    for arch in ("qwen", "flux"):
        forward(ref, arch, man)
    with open(os.path.join(GOLD, "ref_manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    print({k: os.path.getsize(os.path.join(GOLD, k)) for k in os.listdir(GOLD)})


if __name__ == "__main__":
    main()
