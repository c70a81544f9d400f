"""The reference's OWN code driving the drop-in object (CPU; skipped where /root/reference is absent).

INTEGRATION.md section 1 claims that after the import-level edit ``from ragb_vae_b200 import RgbaAutoencoder as
AutoencoderKL`` the reference's ``adapt_vae_to_rgba``, ``_maybe_restore_rgba_convs``, ``RgbaVAE.from_pretrained_rgb``,
``RgbaVAE.forward`` and the ``evaluate_rgba_vae`` loop run unchanged.  tests/refshim.py makes exactly that edit (through
``sys.modules``; the reference files are executed where they lie) and these tests check the claim.

A GPU and ``/root/reference`` never meet (the reference cannot travel to the GPU box, this container has no GPU), so the
arithmetic behind ``RgbaAutoencoder._encode_moments`` / ``_decode_image`` and behind the posterior's ``sample`` / ``kl`` is
replaced HERE, in the test, by the oracle's -- everything else (the diffusers-style surface, config mutation, parameter
replacement, checkpoint I/O, tiling / slicing dispatch, output containers) is the product's host code.  The kernels
themselves meet the same fixtures on the GPU in tests/test_gpu_reference_fixtures.py and the oracle in
tests/test_gpu_parity.py.
"""
import json
import os

import pytest
import torch

import refshim
from oracle import vae_oracle as O

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference is not present on this box")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def R(lib_built):
    import ragb_vae_b200 as R

    return R


@pytest.fixture(scope="module")
def host(R):
    """(CpuBacked RgbaAutoencoder class, CPU posterior class): product host code over oracle arithmetic."""
    from ragb_vae_b200 import autoencoder as A
    from ragb_vae_b200.posterior import DiagonalGaussianDistribution as Posterior

    class CpuPosterior(Posterior):
        def sample(self, generator=None, noise=None, shift=0.0, scale=1.0):
            if noise is None:  # same draw as diffusers' randn_tensor / the product's sample()
                noise = torch.randn(self.mean.shape, generator=generator, dtype=self.parameters.dtype)
            return (self.mean + self.std * noise - shift) * scale

        def kl(self, other=None):
            return O.DiagonalGaussianDistribution(self.parameters).kl(None if other is None else O.DiagonalGaussianDistribution(other.parameters))

    class CpuBacked(R.RgbaAutoencoder):
        def _oracle(self):
            with torch.random.fork_rng():  # module construction draws init values; the caller's RNG stream must not move
                o = O.OracleVAE(self.arch, self.encoder.conv_in.in_channels, self.decoder.conv_out.out_channels)
            o.load_state_dict(self.state_dict())
            return o.eval().requires_grad_(False)

        def _encode_moments(self, x, in_scale=1.0, in_shift=0.0):
            return self._oracle().encode_moments(x * in_scale + in_shift)

        def _decode_image(self, z, out_scale=1.0, out_shift=0.0, clamp=None, z_scale=1.0, z_shift=0.0, model_clamp=True):
            o = self._oracle()
            z = z * z_scale + z_shift
            if self.arch == "qwen" and not model_clamp:
                y = o.decoder(o.post_quant_conv.forward_frame(z))
            else:
                y = o.decode(z).sample
            y = y * out_scale + out_shift
            return y if clamp is None else y.clamp(clamp[0], clamp[1])

    mp = pytest.MonkeyPatch()
    mp.setattr(A, "DiagonalGaussianDistribution", CpuPosterior)
    yield CpuBacked, CpuPosterior
    mp.undo()


@pytest.fixture(scope="module")
def ref(host):
    cls, posterior = host
    return refshim.load_reference(cls, posterior)


def test_reference_adapt_vae_to_rgba_on_the_dropin(ref, host, R):
    cls, _ = host
    for arch in ("flux", "qwen"):
        torch.manual_seed(3)
        a = cls(arch, 3, 3)
        b = cls(arch, 3, 3)
        b.load_state_dict(a.state_dict())
        bias_cache = a._f32(a.decoder.conv_out.bias, "bias")        # a packed copy made before widening ...
        ref.rgba_vae.adapt_vae_to_rgba(a, alpha_bias_init=0.7)      # the reference's function
        R.adapt_vae_to_rgba(b, alpha_bias_init=0.7)                 # the mirror
        assert a.encoder.conv_in.in_channels == 4 and a.decoder.conv_out.out_channels == 4
        assert a.config.in_channels == 4 and a.config.out_channels == 4 and a.config["in_channels"] == 4
        assert a.encoder.conv_in.weight.shape[1] == 4 and torch.all(a.encoder.conv_in.weight[:, 3] == 0)
        assert a.decoder.conv_out.weight.shape[0] == 4 and torch.all(a.decoder.conv_out.weight[3] == 0)
        assert float(a.decoder.conv_out.bias[3]) == pytest.approx(0.7)
        for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
            assert torch.equal(v, w), k
        # ... is not served for the replaced parameter (cache keyed on parameter identity / version)
        assert a._f32(a.decoder.conv_out.bias, "bias").shape[0] == 4 and bias_cache.shape[0] == 3
        ref.rgba_vae.adapt_vae_to_rgba(a)   # idempotent on an RGBA model
        assert float(a.decoder.conv_out.bias[3]) == pytest.approx(0.7)


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_reference_from_pretrained_rgb_loads_a_dropin_checkpoint(ref, host, tmp_path, fmt):
    """rgba_vae.py:231-272: from_pretrained(config says 3, weights say 4, ignore_mismatched_sizes) -> adapt ->
    _maybe_restore_rgba_convs re-reads the 4-channel tensors from the weight file."""
    cls, _ = host
    torch.manual_seed(5)
    src = cls("flux", 4, 4)
    d = tmp_path / "ckpt" / "vae"
    src.save_pretrained(str(d))
    cfg = json.load(open(d / "config.json"))
    cfg["in_channels"] = cfg["out_channels"] = 3      # what diffusers' save_pretrained leaves behind
    json.dump(cfg, open(d / "config.json", "w"))
    if fmt == "bin":
        torch.save({k: v.clone() for k, v in src.state_dict().items()}, d / "pytorch_model.bin")
        os.remove(d / "diffusion_pytorch_model.safetensors")
    with pytest.warns(UserWarning):
        model = ref.rgba_vae.RgbaVAE.from_pretrained_rgb(str(tmp_path / "ckpt"), subfolder="vae", torch_dtype=torch.float32,
                                                         alpha_bias_init=0.3, loss_reduce_mean=True)
    assert isinstance(model.vae, cls) and model.loss_reduce_mean is True
    for k, v in src.state_dict().items():
        assert torch.equal(v, model.vae.state_dict()[k]), k
    # a NaN in the stored RGBA conv must raise like rgba_vae.py:186-191
    bad = {k: v.clone() for k, v in src.state_dict().items()}
    bad["encoder.conv_in.weight"][0, 3, 0, 0] = float("nan")
    if fmt == "bin":
        torch.save(bad, d / "pytorch_model.bin")
    else:
        from safetensors.torch import save_file

        save_file(bad, str(d / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(RuntimeError, match="NaN/Inf"), pytest.warns(UserWarning):
        ref.rgba_vae.RgbaVAE.from_pretrained_rgb(str(tmp_path / "ckpt"), subfolder="vae")


def test_mirror_restore_matches_reference_restore(ref, host, R, tmp_path):
    from ragb_vae_b200.rgba_vae import _maybe_restore_rgba_convs

    cls, _ = host
    torch.manual_seed(6)
    src = cls("qwen", 4, 4)
    src.save_pretrained(str(tmp_path / "q"))
    a, b = cls("qwen", 4, 4), cls("qwen", 4, 4)
    ref.rgba_vae._maybe_restore_rgba_convs(a, str(tmp_path / "q"), None)
    assert _maybe_restore_rgba_convs(b, str(tmp_path / "q"), None) is True
    for name in ("encoder.conv_in.weight", "encoder.conv_in.bias", "decoder.conv_out.weight", "decoder.conv_out.bias"):
        assert torch.equal(a.state_dict()[name], src.state_dict()[name]) and torch.equal(b.state_dict()[name], src.state_dict()[name])
    # nothing to restore: both are silent no-ops
    assert _maybe_restore_rgba_convs(b, str(tmp_path / "missing"), "vae") is False
    ref.rgba_vae._maybe_restore_rgba_convs(a, str(tmp_path / "missing"), "vae")


@pytest.mark.parametrize("arch", ["qwen", "flux"])
def test_reference_forward_and_eval_loop_over_the_dropin(ref, host, arch, oracle_model):
    """The reference's RgbaVAE.forward / evaluate_rgba_vae call encode().latent_dist.sample(), decode().sample,
    next(model.parameters()).dtype ... on the drop-in and reproduce the fixtures made with the oracle VAE."""
    from types import SimpleNamespace

    from safetensors.torch import load_file

    cls, posterior = host
    g = load_file(os.path.join(GOLD, f"ref_forward_{arch}.safetensors"))
    man = json.load(open(os.path.join(GOLD, "ref_manifest.json")))
    entry = [e for e in man["forward"] if e["arch"] == arch][0]
    vae = cls(arch)
    vae.load_state_dict(oracle_model(arch).state_dict())
    vae.enable_slicing()            # rgba_vae_stage.py:300-304 (numerically a no-op)
    vae.enable_tiling()             # :296-299; 64x64 is below every tile size
    model = ref.rgba_vae.RgbaVAE(vae=vae)
    torch.manual_seed(42)
    with torch.no_grad():
        recon, post = model(g["x"])
        recon3, _ = model(g["x"][:1, :3])
    assert isinstance(post, posterior)
    assert torch.allclose(recon, g["recon"], atol=1e-5) and torch.allclose(post.parameters, g["moments"], atol=1e-4)
    assert torch.allclose(recon3, g["recon3"], atol=1e-5)
    lines = []
    acc = SimpleNamespace(device=torch.device("cpu"), gather=lambda v: v, print=lambda s: lines.append(s), is_main_process=False)
    torch.manual_seed(45)
    ref.stage.evaluate_rgba_vae(acc, model, [{"composite": g["eval_batch0"]}, {"composite": g["eval_batch1"]}], epoch=3,
                                eval_cfg={"val_background_colors": ["white", "black", (0.2, 0.5, 0.9)]})
    assert lines == entry["eval_lines"]
