"""ctypes binding of librgbavae.so (include/rgbavae.h).  The library is the only compute path:
if it is missing or a call fails, this module raises -- there is no eager/CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "librgbavae.so")

RV_F32, RV_BF16 = 0, 1
ABI_VERSION = 25
PROF_CATEGORIES = 10
PROF_NAMES = ("conv_tc", "conv_direct", "norm_silu", "softmax", "layout", "reparam", "recon_loss", "composite_psnr",
              "attention", "conv_tc_upsample")


class RvError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad_lo", C.c_int32), ("upsample", C.c_int32),
        ("oh", C.c_int32), ("ow", C.c_int32),
        ("x_dtype", C.c_int32), ("y_dtype", C.c_int32),
        ("x_nchw", C.c_int32), ("y_nchw", C.c_int32),
        ("x_cstride", C.c_int32), ("y_cstride", C.c_int32),
        ("bias_mode", C.c_int32),
        ("in_scale", C.c_float), ("in_shift", C.c_float),
        ("out_scale", C.c_float), ("out_shift", C.c_float),
        ("clamp", C.c_int32), ("clamp_lo", C.c_float), ("clamp_hi", C.c_float),
        ("alpha", C.c_float),
        ("taps_1d", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float

# name -> (restype, argtypes); mirrors include/rgbavae.h one to one
SIGNATURES = {
    "rv_abi_version": (_I, []),
    "rv_last_error": (C.c_char_p, []),
    "rv_init": (_I, []),
    "rv_launch_count": (_L, []),
    "rv_prof_begin": (_I, []),
    "rv_prof_end": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "rv_conv2d_direct": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "rv_conv2d_tc": (_I, [C.POINTER(ConvDesc), _P, _P, _L, _P, _P, _P, _P]),
    "rv_conv_out": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "rv_conv2d_tc_norm": (_I, [C.POINTER(ConvDesc), _P, _P, _L, _P, _P, _P, _P, _P, _I, _P]),
    "rv_conv2d_tc_gnstats_scratch_bytes": (_L, [C.POINTER(ConvDesc), _I]),
    "rv_conv2d_tc_gnstats": (_I, [C.POINTER(ConvDesc), _P, _P, _L, _P, _P, _P, _I, _P, _P, _L, _P]),
    "rv_pack_conv_weights": (_I, [_P, _I, _I, _I, _I, _P, C.POINTER(C.c_int64), _P]),
    "rv_pack_conv_weights_direct": (_I, [_P, _I, _I, _I, _P, _P]),
    "rv_rmsnorm_silu": (_I, [_P, _P, _P, _L, _I, _I, _I, _P]),
    "rv_groupnorm_stats": (_I, [_P, _P, _I, _L, _I, _I, _I, _P]),
    "rv_groupnorm_silu": (_I, [_P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _I, _P]),
    "rv_softmax_rows": (_I, [_P, _P, _L, _L, _L, _L, _I, _P]),
    "rv_attention": (_I, [_P, _P, _L, _P, _P, _L, _I, _I, _I, _P]),
    "rv_attention_lse": (_I, [_P, _P, _L, _P, _P, _L, _P, _I, _I, _I, _P]),
    "rv_attention_workspace_bytes": (_L, [_I, _I]),
    "rv_attention_ws": (_I, [_P, _P, _L, _P, _L, _L, _P, _L, _P, _P, _L, _I, _I, _I, _P]),
    "rv_gemm_rowstat": (_I, [C.POINTER(ConvDesc), _P, _P, _L, _P, _I, _P, _P, _P]),
    "rv_rowdot": (_I, [_P, _P, _L, _I, _L, _L, _F, _P, _P]),
    "rv_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _L, _I, _I, _I, _F, _F, _P]),
    "rv_nchw_to_nhwc_hpack": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _P]),
    "rv_nhwc_to_nchw": (_I, [_P, _P, _I, _I, _L, _I, _I, _I, _P]),
    "rv_triplet_augment": (_I, [_P, _P, _I, _L, _I, _P]),
    "rv_background_blend": (_I, [_P, _P, _P, _P, _I, _L, _I, _P]),
    "rv_pack_latents": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _P]),
    "rv_blend_tiles": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rv_rgba_u8_to_nchw": (_I, [_P, _P, _I, _L, _I, _F, _F, _P]),
    "rv_nchw_to_rgba_u8": (_I, [_P, _P, _I, _L, _I, _P]),
    "rv_reparam": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _F, _P]),
    "rv_recon_loss": (_I, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _P, _P, _I, _L, _I, _P]),
    "rv_composite_psnr": (_I, [_P, _P, C.POINTER(C.c_float), _I, _P, _P, _I, _L, _I, _P]),
    "rv_rgba_loss_terms": (_I, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, _P, _I, _L, _I, _P]),
    "rv_reduce_blocks": (_I, [_L]),
    "rv_recon_loss_bwd": (_I, [_P, _P, C.POINTER(C.c_float), C.POINTER(C.c_float), _I, _F, _F, _F, _P, _I, _L, _I, _P]),
    "rv_reparam_bwd": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _P]),
    "rv_rmsnorm_silu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _F, _L, _I, _I, _I, _P]),
    "rv_conv2d_wgrad": (_I, [_P, _P, _P, _L, _L, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "rv_pack_dgrad_weights": (_I, [_P, _L, _L, _P, _I, _I, _I, _I, _P]),
    "rv_resample2x": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "rv_add_bf16": (_I, [_P, _P, _P, _L, _P]),
    "rv_softmax_bwd": (_I, [_P, _P, _P, _P, _L, _L, _L, _L, _F, _P]),
    "rv_grad_sqnorm": (_I, [_P, _L, _P, _P, _P]),
    "rv_grad_sqnorm_scratch_bytes": (_I, []),
    "rv_adamw_step": (_I, [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P, _F, _P, _F, _P]),
    "rv_adamw_advance": (_I, [_P, _F, _F, _P]),
    "rv_groupnorm_silu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _I, _I, _P]),
    "rv_kl_ref": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the library (once) and binds every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RvError(
            f"{LIB_PATH} is missing: build it with `python ragb_vae_b200/build.py` (needs nvcc). "
            "ragb_vae_b200 has no CPU or eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    got = lib.rv_abi_version()
    if got != ABI_VERSION:
        raise RvError(f"librgbavae ABI {got} != expected {ABI_VERSION}; rebuild with `python ragb_vae_b200/build.py --force`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rv_last_error().decode("utf-8", "replace")
        raise RvError(f"{what} failed (code {rc}): {msg}")
