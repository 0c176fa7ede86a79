"""ctypes binding of libtheatergen.so (the C-ABI in include/theatergen.h).

The product path fails loudly when the library is missing -- there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtheatergen.so")
_lib = None

vp, ci, cf = C.c_void_p, C.c_int, C.c_float

_SIGS = {
    "tg_last_error": (C.c_char_p, []),
    "tg_version": (ci, []),
    "tg_launch_count": (C.c_ulonglong, []),
    "tg_init": (ci, []),
    "tg_gemm": (ci, [vp, ci, ci, vp, ci, vp, ci, ci, ci, vp, vp, ci, ci, vp, ci, ci, cf, vp, ci, vp]),
    "tg_conv3x3": (ci, [vp, ci, ci, ci, ci, vp, ci, vp, vp, ci, vp, ci, ci, cf, vp, ci, vp]),
    "tg_groupnorm": (ci, [vp, ci, vp, ci, ci, ci, vp, vp, cf, ci, vp, vp, vp]),
    "tg_gemm_set_prof": (ci, [vp]),
    "tg_attn_set_prof": (ci, [vp]),
    "tg_attn_set_variant": (ci, [ci]),
    "tg_attn_cross_set_impl": (ci, [ci]),
    "tg_conv3x3_s2": (ci, [vp, ci, ci, ci, ci, vp, ci, vp, ci, ci, vp, ci, vp]),
    "tg_upsample2x_conv3x3": (ci, [vp, ci, ci, ci, ci, vp, ci, vp, vp, vp]),
    "tg_gemm_set_scratch_slot": (ci, [ci]),
    "tg_set_pdl": (ci, [ci]),
    "tg_cross_scores": (ci, [vp, ci, vp, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp]),
    "tg_cross_apply": (ci, [vp, ci, vp, ci, ci, ci, ci, vp, vp, ci, vp, ci, vp]),
    "tg_gemm_set_stats": (ci, [vp, ci]),
    "tg_groupnorm_apply": (ci, [vp, ci, vp, vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, cf, ci, vp, vp]),
    "tg_groupnorm_tune": (ci, [ci, ci, ci]),
    "tg_layernorm": (ci, [vp, ci, vp, vp, cf, vp, ci, ci, ci, vp]),
    "tg_attention": (ci, [vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, ci, cf, ci, vp]),
    "tg_attention_tc": (ci, [vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, cf, vp]),
    "tg_attention_tc_packed": (ci, [vp, vp, vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, cf, vp]),
    "tg_attention_cross_ip": (ci, [vp, ci, vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, cf, vp]),
    "tg_upsample2x": (ci, [vp, ci, ci, ci, ci, vp, vp]),
    "tg_im2col_s2": (ci, [vp, ci, ci, ci, ci, vp, ci, vp]),
    "tg_im2col_patch": (ci, [vp, vp, ci, ci, ci, ci, ci, ci, vp]),
    "tg_embed_tokens": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "tg_nhwc_head_to_nchw": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "tg_gemm_set_tuned": (ci, [ci, ci, ci, ci, ci, ci, ci, ci]),
    "tg_gemm_clear_tuned": (ci, []),
    "tg_gemm_force_tile": (ci, [ci, ci, ci]),
    "tg_layernorm_bwd": (ci, [vp, ci, vp, vp, ci, vp, ci, ci, ci, cf, vp]),
    "tg_groupnorm_bwd": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, cf, ci, vp]),
    "tg_softmax_bwd_rows": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, cf, vp]),
    "tg_geglu_bwd": (ci, [vp, vp, vp, ci, ci, vp]),
    "tg_attention_cross_ip_bwd_q": (ci, [vp, ci, vp, vp, vp, vp, ci, ci, ci, cf, vp, ci, vp, vp, ci, ci, ci, ci, cf, vp]),
    "tg_upsample2x_bwd": (ci, [vp, ci, ci, ci, ci, vp, vp]),
    "tg_col2im_s2": (ci, [vp, ci, ci, ci, ci, vp, vp]),
    "tg_guidance_update": (ci, [vp, vp, cf, cf, vp, vp, ci, vp]),
    "tg_gaussian_sample": (ci, [vp, vp, vp, ci, ci, ci, cf, vp]),
    "tg_conv_in": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]),
    "tg_im2col_small": (ci, [vp, vp, ci, ci, ci, ci, vp]),
    "tg_conv_out": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]),
    "tg_pointwise_nchw": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, vp]),
    "tg_softmax_rows": (ci, [vp, ci, vp, ci, ci, ci, cf, vp]),
    "tg_timestep_embedding": (ci, [vp, vp, ci, ci, vp]),
    "tg_add_inplace": (ci, [vp, vp, C.c_longlong, vp]),
    "tg_nchw_to_nhwc": (ci, [vp, vp, ci, ci, ci, vp]),
    "tg_duplicate_latents": (ci, [vp, vp, ci, vp]),
    "tg_cfg_ddim_composite": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, vp, ci, vp]),
    "tg_guidance_energy": (ci, [vp, vp, vp, ci, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci, cf, cf, cf, cf, cf, cf,
                                vp, vp, vp]),
}


class GuidanceItem(C.Structure):
    _fields_ = [("key", C.c_ubyte), ("obj", C.c_ubyte), ("box", C.c_ubyte), ("ref", C.c_ubyte),
                ("tok", C.c_int), ("weight", C.c_float)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m theatergen_b200.build` "
                "(theatergen_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
        _load_tuned(l)
    return _lib


TUNED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuned.json")


def _load_tuned(l):
    """Measured GEMM tile choices per problem shape (tools/gemm_autotune.py); TG_GEMM_TUNED=0 keeps the cost model only."""
    if os.environ.get("TG_GEMM_TUNED", "1") == "0" or not os.path.exists(TUNED_PATH):
        return 0
    import json
    n = 0
    for M, N, K, act, conv, bn, split, pair in json.load(open(TUNED_PATH))["entries"]:
        if l.tg_gemm_set_tuned(M, N, K, act, conv, bn, split, pair) == 0:
            n += 1
    return n


def exported_symbols():
    return list(_SIGS.keys())


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().tg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"theatergen {what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(lib().tg_launch_count())
