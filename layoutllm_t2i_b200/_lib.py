"""ctypes binding of libltt_b200.so (the C-ABI declared in include/ltt_b200.h).

The library is the product: there is no Python/PyTorch fallback.  Import fails loudly when the shared object is
missing (build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``layoutllm_t2i_b200/csrc/build.sh``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LTT_LIB: development aid for A/B-ing two builds of the library on the GPU box (default: the in-tree build)
LIB_PATH = os.environ.get("LTT_LIB") or os.path.join(_HERE, "libltt_b200.so")


class LttError(RuntimeError):
    pass


class UNetConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("model_channels", C.c_int),
                ("num_res_blocks", C.c_int), ("n_levels", C.c_int), ("channel_mult", C.c_int * 8),
                ("n_attn_res", C.c_int), ("attention_resolutions", C.c_int * 8), ("num_heads", C.c_int),
                ("context_dim", C.c_int), ("grounding_in_dim", C.c_int), ("grounding_out_dim", C.c_int),
                ("fourier_freqs", C.c_int), ("max_objs", C.c_int)]


class VaeConfig(C.Structure):
    _fields_ = [("ch", C.c_int), ("out_ch", C.c_int), ("n_levels", C.c_int), ("ch_mult", C.c_int * 8),
                ("num_res_blocks", C.c_int), ("z_channels", C.c_int), ("embed_dim", C.c_int), ("scale_factor", C.c_float)]


class ClipConfig(C.Structure):
    _fields_ = [("vocab", C.c_int), ("max_pos", C.c_int), ("hidden", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
                ("ffn", C.c_int), ("eps", C.c_float), ("act", C.c_int), ("proj_dim", C.c_int), ("eos_token_id", C.c_int)]


class ClipVisionConfig(C.Structure):
    _fields_ = [("image_size", C.c_int), ("patch", C.c_int), ("hidden", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
                ("ffn", C.c_int), ("eps", C.c_float), ("act", C.c_int), ("proj_dim", C.c_int)]


_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); mirrors include/ltt_b200.h one to one
_SIGS = {
    "ltt_last_error": (C.c_char_p, []),
    "ltt_version": (C.c_char_p, []),
    "ltt_create": (_i, [C.POINTER(UNetConfig), _i, C.POINTER(_vp)]),
    "ltt_destroy": (None, [_vp]),
    "ltt_load_param": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "ltt_finalize": (_i, [_vp]),
    "ltt_set_first_conv": (_i, [_vp, _vp, _vp, _i]),
    "ltt_clear_first_conv": (_i, [_vp]),
    "ltt_set_conditioning": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ltt_unet_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "ltt_plms_sample": (_i, [_vp, _vp, _i, _i, C.POINTER(_i), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f),
                             C.POINTER(_f), _f, _vp, _vp, _vp]),
    "ltt_launch_count": (_i64, [_vp]),
    "ltt_profile_enable": (_i, [_vp, _i]),
    "ltt_profile_report": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "ltt_debug_set_taps": (_i, [_vp, _vp, _i64]),
    "ltt_debug_tap_count": (_i, [_vp]),
    "ltt_debug_tap_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "ltt_vae_create": (_i, [C.POINTER(VaeConfig), _i, C.POINTER(_vp)]),
    "ltt_vae_destroy": (None, [_vp]),
    "ltt_vae_load_param": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "ltt_vae_finalize": (_i, [_vp]),
    "ltt_vae_decode": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "ltt_vae_launch_count": (_i64, [_vp]),
    "ltt_clip_create": (_i, [C.POINTER(ClipConfig), _i, C.POINTER(_vp)]),
    "ltt_clip_destroy": (None, [_vp]),
    "ltt_clip_load_param": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "ltt_clip_finalize": (_i, [_vp]),
    "ltt_clip_encode": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "ltt_clip_launch_count": (_i64, [_vp]),
    "ltt_clip_vision_create": (_i, [C.POINTER(ClipVisionConfig), _i, C.POINTER(_vp)]),
    "ltt_clip_vision_destroy": (None, [_vp]),
    "ltt_clip_vision_load_param": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "ltt_clip_vision_finalize": (_i, [_vp]),
    "ltt_clip_vision_encode": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ltt_clip_vision_preprocess": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp]),
    "ltt_clip_vision_launch_count": (_i64, [_vp]),
    "ltt_reward_head": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), _vp, _vp, _vp, _vp, _vp, _vp]),
    "ltt_op_linear": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _f, _i, _vp, _i, _i, _vp]),
    "ltt_op_linear_ln_linear": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _f, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "ltt_op_pack_geglu": (_i, [_vp, _i, _i, _vp, _vp]),
    "ltt_op_pack_conv3x3": (_i, [_vp, _i, _i, _vp, _vp]),
    "ltt_op_conv3x3": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "ltt_op_qkv": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "ltt_op_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp]),
    "ltt_op_attention_causal": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp]),
    "ltt_op_groupnorm": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _i, _vp, _vp]),
    "ltt_op_layernorm": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp]),
    "ltt_op_rela_rects": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ltt_op_rela_pool": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "ltt_op_rela_scatter": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "ltt_op_rela_scatter_ln": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp]),
    "ltt_op_rela_fold": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "ltt_op_rela_attn_fused": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "ltt_op_small_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "ltt_op_posnet_input": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ltt_op_timestep_embedding": (_i, [_vp, _i, _i, _vp, _vp]),
    "ltt_op_plms_update": (_i, [_vp, _vp, _f, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _i64, _vp]),
}

EXPORTS = tuple(_SIGS)


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise LttError(f"{LIB_PATH} is missing: the sm_100a extension must be built (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    partial = bool(os.environ.get("LTT_LIB")) and os.environ.get("LTT_LIB_PARTIAL") == "1"      # A/B against an OLDER build only
    for name, (res, args) in _SIGS.items():
        if partial and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ltt_last_error().decode(errors="replace")
        raise LttError(f"{what} failed ({rc}): {msg}")


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor, None -> NULL."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
