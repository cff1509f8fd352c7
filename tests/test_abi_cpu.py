"""The C-ABI library loads without a GPU and exports exactly what include/ltt_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ltt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltt_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from layoutllm_t2i_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_library_loads_and_exports_every_symbol():
    from layoutllm_t2i_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()          # getattr on every declared symbol: AttributeError if one is missing
    assert lib.ltt_version().decode().startswith("ltt_b200")
    assert lib.ltt_launch_count(None) == 0
    assert lib.ltt_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch):
    from layoutllm_t2i_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(_lib.LttError, match="no CPU fallback"):
        _lib.load()


def test_engine_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from layoutllm_t2i_b200 import _lib
    from layoutllm_t2i_b200.engine import Engine
    with pytest.raises(_lib.LttError, match="no CPU fallback"):
        Engine({}, 0)


def test_no_oracle_import_in_product_package():
    pkg = os.path.join(ROOT, "layoutllm_t2i_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(d, f)


def test_side_model_handles_need_cuda():
    """The conditioning-prep / reward / VAE wrappers refuse to run without a CUDA device (no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from layoutllm_t2i_b200 import _lib
    from layoutllm_t2i_b200.clip import ClipTextEncoder, ClipVisionEncoder
    from layoutllm_t2i_b200.vae import VaeDecoder
    for ctor in (lambda: ClipTextEncoder({}, 0), lambda: ClipVisionEncoder({}, 0), lambda: VaeDecoder({}, 0)):
        with pytest.raises(_lib.LttError, match="no CPU fallback"):
            ctor()


def test_argument_validation_needs_no_gpu():
    """Bad configurations / null arguments are rejected with an error code and a message before any CUDA call."""
    import ctypes as C
    from layoutllm_t2i_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    bad_text = _lib.ClipConfig(vocab=1000, max_pos=77, hidden=100, heads=2, layers=1, ffn=256, eps=1e-5, act=0, proj_dim=0, eos_token_id=2)
    assert lib.ltt_clip_create(C.byref(bad_text), 0, C.byref(h)) == -1 and b"heads of 64" in lib.ltt_last_error()
    bad_act = _lib.ClipConfig(vocab=1000, max_pos=77, hidden=128, heads=2, layers=1, ffn=256, eps=1e-5, act=1, proj_dim=0, eos_token_id=2)
    assert lib.ltt_clip_create(C.byref(bad_act), 0, C.byref(h)) == -1
    bad_vis = _lib.ClipVisionConfig(image_size=225, patch=14, hidden=128, heads=2, layers=1, ffn=256, eps=1e-5, act=0, proj_dim=0)
    assert lib.ltt_clip_vision_create(C.byref(bad_vis), 0, C.byref(h)) == -1 and b"vision tower" in lib.ltt_last_error()
    assert lib.ltt_clip_create(None, 0, C.byref(h)) == -1 and lib.ltt_clip_vision_create(None, 0, None) == -1
    assert lib.ltt_clip_encode(None, None, 1, 77, None, None, None, None) == -8
    assert lib.ltt_clip_vision_encode(None, None, 1, None, None, None, None) == -8
    assert lib.ltt_clip_vision_preprocess(None, None, 1, 8, 8, None, None, None, None) == -1
    assert lib.ltt_reward_head(None, None, None, 1, 768, None, None, None, None, None, None, None, None, None) == -1
    assert b"ltt_reward_head" in lib.ltt_last_error()
    assert lib.ltt_clip_launch_count(None) == 0 and lib.ltt_clip_vision_launch_count(None) == 0 and lib.ltt_vae_launch_count(None) == 0
    for mult in ((1, 2), (0, 0)):           # channels not a multiple of 64; no channels at all
        bad_vae = _lib.VaeConfig(ch=100, out_ch=3, n_levels=2, num_res_blocks=1, z_channels=4, embed_dim=4, scale_factor=0.18215)
        bad_vae.ch_mult[0], bad_vae.ch_mult[1] = mult
        assert lib.ltt_vae_create(C.byref(bad_vae), 0, C.byref(h)) == -1, mult
