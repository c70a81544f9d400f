"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the model's parameters/keys match the oracle (= diffusers) layout, checkpoint I/O
follows the reference's "config says 3, weights say 4" rule, errors mirror the reference's.
No kernel is launched here (no GPU in this container)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "rgbavae.h")).read()
    declared = set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(lib_built)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in rgbavae.h but not exported"
    from ragb_vae_b200 import _lib

    assert set(_lib.SIGNATURES) == declared, "python binding table and header disagree"
    assert lib.rv_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.ConvDesc) == 27 * 4


def test_library_is_blackwell_native(lib_built):
    """SASS evidence that the conv kernel is tcgen05 + TMA (B200_PROFILING.md table)."""
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path must not be present"
    elf = subprocess.run(["cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in elf


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_state_dict_layout_matches_diffusers_names(arch, lib_built, oracle_model):
    import ragb_vae_b200 as R

    m = R.RgbaAutoencoder(arch)
    o = oracle_model(arch)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so) == set(sm)
    assert all(so[k].shape == sm[k].shape for k in so)
    m.load_state_dict(so)
    assert m.config.in_channels == 4 and m.encoder.conv_in.in_channels == 4 and m.decoder.conv_out.out_channels == 4
    if arch == "qwen":
        assert m.encoder.conv_in.weight.dim() == 5 and m.encoder.conv_in.weight2d().shape == (96, 4, 3, 3)
        assert torch.equal(m.encoder.conv_in.weight2d(), so["encoder.conv_in.weight"][:, :, 2])
        assert m.config.block_out_channels == [96, 192, 384, 384] and m.config.sample_size == 256
    else:
        assert m.config.scaling_factor == 0.3611 and m.config.shift_factor == 0.1159 and m.config.sample_size == 1024


def test_checkpoint_roundtrip_and_config3_weights4_rule(tmp_path, lib_built):
    import ragb_vae_b200 as R
    from ragb_vae_b200.rgba_vae import _maybe_restore_rgba_convs

    torch.manual_seed(0)
    rgb = R.RgbaAutoencoder("flux", 3, 3)
    R.adapt_vae_to_rgba(rgb, alpha_bias_init=0.25)
    assert rgb.encoder.conv_in.weight.shape[1] == 4 and torch.all(rgb.encoder.conv_in.weight[:, 3] == 0)
    assert float(rgb.decoder.conv_out.bias[3]) == 0.25
    with torch.no_grad():  # make the alpha tensors recognisable
        rgb.encoder.conv_in.weight[:, 3] = 0.5
    d = tmp_path / "ckpt"
    rgb.save_pretrained(str(d))
    # emulate diffusers: the saved config still says 3 channels
    import json

    cfg = json.load(open(d / "config.json"))
    cfg["in_channels"] = cfg["out_channels"] = 3
    json.dump(cfg, open(d / "config.json", "w"))
    with pytest.raises(RuntimeError):
        R.RgbaAutoencoder.from_pretrained(str(d))
    with pytest.warns(UserWarning):
        re3 = R.RgbaAutoencoder.from_pretrained(str(d), ignore_mismatched_sizes=True)
    assert re3.encoder.conv_in.weight.shape[1] == 3
    R.adapt_vae_to_rgba(re3)
    assert _maybe_restore_rgba_convs(re3, str(d), None)
    assert torch.equal(re3.encoder.conv_in.weight, rgb.encoder.conv_in.weight)
    assert torch.equal(re3.decoder.conv_out.bias, rgb.decoder.conv_out.bias)
    # the one-call path the reference uses
    wrapped = R.RgbaVAE.from_pretrained_rgb(str(d), subfolder="vae")
    assert torch.equal(wrapped.vae.encoder.conv_in.weight, rgb.encoder.conv_in.weight)
    for k, v in rgb.state_dict().items():
        assert torch.equal(v, wrapped.vae.state_dict()[k]), k


def test_errors_mirror_reference(lib_built):
    import ragb_vae_b200 as R
    from ragb_vae_b200._lib import RvError
    from ragb_vae_b200.rgba_vae import _ensure_alpha

    with pytest.raises(ValueError):
        R.RgbaAutoencoder("sdxl")
    with pytest.raises(ValueError):
        R.AlphaVaeLoss(custom_eb=(1.0, 2.0))
    with pytest.raises(ImportError):
        R.AlphaVaeLoss(use_lpips=True)
    x = torch.rand(1, 4, 8, 8)
    with pytest.raises(ValueError):
        R.composite_over_background(x, (1.0, 0.0))
    with pytest.raises(ValueError):
        R.composite_over_background(x, torch.zeros(3, 4, 4))
    assert _ensure_alpha(x[:, :3]).shape[1] == 4 and torch.all(_ensure_alpha(x[:, :3])[:, 3] == 1)
    vae = R.RgbaAutoencoder("qwen")
    with pytest.raises(ValueError):
        vae.encode(torch.rand(1, 4, 30, 32))
    with pytest.raises(ValueError):
        vae.encode(torch.rand(1, 3, 32, 32))
    with pytest.raises(RvError):  # CPU tensors: the product path refuses, it never falls back
        vae.encode(torch.rand(1, 4, 32, 32))
    with pytest.raises(ValueError):
        R.DiagonalGaussianDistribution(torch.zeros(1, 3, 2, 2))


def test_product_package_does_not_import_the_oracle(lib_built):
    pkg = os.path.join(ROOT, "ragb_vae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
