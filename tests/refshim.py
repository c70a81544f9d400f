"""Loads the UNMODIFIED reference sources (``/root/reference/src/models/{losses,rgba_vae}.py`` and
``src/training/rgba_vae_stage.py``) by file path behind a ``sys.modules`` shim (TEST INFRASTRUCTURE ONLY).

The reference imports ``diffusers`` (not installable here), ``accelerate``, ``matplotlib`` ... at module
top.  Only two names of those packages matter on the RGBA-VAE path: ``diffusers.AutoencoderKL`` (used for
``AutoencoderKL.from_pretrained`` and as a type annotation, src/models/rgba_vae.py:17,249) and
``diffusers.models.autoencoders.vae.DiagonalGaussianDistribution`` (losses.py:7, rgba_vae_stage.py:14,700).
The caller chooses what they are bound to:

* the oracle's classes   -> the reference's own functions run on the CPU and produce the committed fixtures
                            (scripts/make_reference_fixtures.py);
* ``ragb_vae_b200``'s     -> the reference's own ``RgbaVAE`` / ``adapt_vae_to_rgba`` / ``_maybe_restore_rgba_convs`` /
                            ``from_pretrained_rgb`` drive the drop-in object (tests/test_reference_dropin.py), which
                            is exactly the import-level edit INTEGRATION.md section 1 describes.

Nothing is copied: the files are executed where they lie.  ``/root/reference`` does not exist on the GPU box,
so everything that uses this module is a CPU test that skips when the directory is absent; the GPU tests use
the fixtures.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("RGBAVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "rgba_vae.py"))


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless placeholder class (packages the RGBA-VAE path never calls:
    accelerate, matplotlib, the dataset modules ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        placeholder = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, placeholder)
        return placeholder


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # looks like a package so that dotted imports resolve through sys.modules
    return m


def load_reference(autoencoder_cls, gaussian_cls) -> SimpleNamespace:
    """Execute the three reference files with ``diffusers.AutoencoderKL`` -> ``autoencoder_cls`` and
    ``DiagonalGaussianDistribution`` -> ``gaussian_cls``.  Returns ``SimpleNamespace(losses, rgba_vae, stage)``
    (the loaded modules).  ``sys.modules`` is restored afterwards."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not present")
    shim = {
        "diffusers": _module("diffusers", AutoencoderKL=autoencoder_cls),
        "diffusers.models": _module("diffusers.models"),
        "diffusers.models.autoencoders": _module("diffusers.models.autoencoders"),
        "diffusers.models.autoencoders.vae": _module("diffusers.models.autoencoders.vae",
                                                      DiagonalGaussianDistribution=gaussian_cls),
    }
    for name in ("accelerate", "accelerate.utils", "matplotlib", "matplotlib.pyplot", "src.data",
                 "src.data.multilayer_dataset", "src.data_generation", "src.data_generation.bucket_dataset"):
        shim[name] = _Anything(name)
    shim["matplotlib"].pyplot = shim["matplotlib.pyplot"]
    shim["src"] = _module("src")
    shim["src.models"] = _module("src.models")
    shim["src.training"] = _module("src.training")
    saved = {k: sys.modules.get(k) for k in list(shim) + ["src.models.losses", "src.models.rgba_vae",
                                                           "src.training.rgba_vae_stage"]}
    sys.modules.update(shim)

    def load(modname: str, relpath: str):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    try:
        losses = load("src.models.losses", "src/models/losses.py")
        rgba_vae = load("src.models.rgba_vae", "src/models/rgba_vae.py")
        # what src/models/__init__.py:6-7 re-exports (its flux_kontext import needs peft and is off the path)
        for name in ("RgbaVAE", "composite_over_background", "composite_over_black", "composite_over_white"):
            setattr(shim["src.models"], name, getattr(rgba_vae, name))
        shim["src.models"].AlphaVaeLoss = losses.AlphaVaeLoss
        stage = load("src.training.rgba_vae_stage", "src/training/rgba_vae_stage.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return SimpleNamespace(losses=losses, rgba_vae=rgba_vae, stage=stage)
