"""ragb_vae_b200 -- B200-native (sm_100a) implementation of the RGBA-VAE hot path of
jaejung-dev/ragb-vae: encode -> reparameterize -> decode, the AlphaVAE reconstruction loss and
the alpha-over-background composite + PSNR validation, behind the diffusers-style surface the
reference calls.  All arithmetic runs in librgbavae.so (hand-written CUDA, C ABI in
include/rgbavae.h); importing this package fails loudly if the library is not built."""
from . import _lib

_lib.load()  # no library, no package: there is no eager / CPU fallback

from .autoencoder import RgbaAutoencoder  # noqa: E402
from .losses import AlphaVaeLoss  # noqa: E402
from .plumbing import (RandomBackgroundBlend, build_detail_augmented_triplet, build_training_batch, pack_latents,  # noqa: E402
                       rgba_u8_to_tensor, split_triplet_distribution, tensor_to_rgba_u8, unpack_latents)
from .posterior import DiagonalGaussianDistribution  # noqa: E402
from .rgba_vae import (RgbaVAE, adapt_vae_to_rgba, composite_over_background, composite_over_black,  # noqa: E402
                       composite_over_white)
from .trainer import VaeTrainStep  # noqa: E402
from .validation import compute_psnr, evaluate_rgba_vae, validation_metrics  # noqa: E402

__all__ = ["RgbaAutoencoder", "RgbaVAE", "AlphaVaeLoss", "DiagonalGaussianDistribution", "adapt_vae_to_rgba",
           "composite_over_background", "composite_over_white", "composite_over_black", "compute_psnr",
           "validation_metrics", "evaluate_rgba_vae", "build_detail_augmented_triplet", "split_triplet_distribution",
           "pack_latents", "unpack_latents", "rgba_u8_to_tensor", "tensor_to_rgba_u8", "VaeTrainStep", "RandomBackgroundBlend",
           "build_training_batch"]
