"""Host-side logic of the drop-in `ldm` tree (no GPU): state_dict grammar, schedule tables, config instantiation,
null grounding input, loud failure without CUDA."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "layoutllm_t2i_b200", "dropin")
if DROPIN not in sys.path:
    sys.path.insert(0, DROPIN)

from oracle import plms_oracle as po  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

TINY = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
            num_res_blocks=1, channel_mult=[1, 2], num_heads=8, transformer_depth=1, context_dim=768,
            fuser_type="gatedSA", grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)


def make_unet(cfg):
    from ldm.util import instantiate_from_config
    return instantiate_from_config(dict(
        target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
        params=dict(image_size=cfg["image_size"], in_channels=4, out_channels=4, model_channels=cfg["model_channels"],
                    attention_resolutions=cfg["attention_resolutions"], num_res_blocks=cfg["num_res_blocks"],
                    channel_mult=cfg["channel_mult"], num_heads=8, transformer_depth=1, context_dim=768,
                    fuser_type="gatedSA", use_checkpoint=True,
                    grounding_tokenizer=dict(target="ldm.modules.diffusionmodules.text_grounding_net.PositionNet",
                                             params=dict(in_dim=768, out_dim=768)))))


def test_state_dict_grammar_matches_reference_spec():
    m = make_unet(TINY)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    want = {k: tuple(s) for k, s, _ in uo.state_dict_spec(TINY)}     # pinned to the reference in tests/gen_golden.py
    assert got == want
    m.load_state_dict(uo.synthetic_state_dict(TINY, seed=7), strict=True)
    assert m._engine_stale


def test_full_config_key_count():
    # full-size grammar without allocating 5 GB: meta device
    with torch.device("meta"):
        m = make_unet(dict(uo.default_unet_config()))
    assert len(m.state_dict()) == 1238
    assert sum(p.numel() for p in m.state_dict().values()) == 1_261_457_796


def test_type_scan_and_scale_write():
    from ldm.modules.attention import GatedCrossAttentionDense, GatedSelfAttentionDense, RelationCrossAttention
    m = make_unet(TINY)
    n = 0
    for mod in m.modules():                      # set_alpha_scale of the reference (txt2img.py:46-50)
        if type(mod) == GatedCrossAttentionDense or type(mod) == GatedSelfAttentionDense:
            mod.scale = 0.25
            n += 1
    assert n == 7 and m.fuser_scale() == 0.25
    assert all(r.scale == 1 for r in m.modules() if type(r) == RelationCrossAttention)


def test_forward_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = make_unet(TINY)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(dict(x=torch.zeros(1, 4, 16, 16), timesteps=torch.zeros(1), context=torch.zeros(1, 77, 768),
               relations=torch.zeros(1, 10, 768)))


def test_plms_schedule_tables_match_reference_golden():
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    g = torch.load(os.path.join(ROOT, "tests", "golden", "plms.pt"), weights_only=False)
    diff = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000)
    assert torch.equal(diff.alphas_cumprod, g["alphas_cumprod"])
    s = PLMSSampler(diff, model=None)
    s.make_schedule(50)
    assert np.array_equal(s.ddim_timesteps, g["ddim_timesteps"].numpy())
    assert torch.equal(s.ddim_alphas, g["ddim_alphas"])
    assert np.array_equal(s.ddim_alphas_prev, g["ddim_alphas_prev"].numpy())
    assert np.allclose(np.asarray(s.ddim_sqrt_one_minus_alphas), g["ddim_sqrt_one_minus_alphas"].numpy(), rtol=1e-7)
    ts, a_t, a_prev, s1m = po.plms_tables(50, diff.alphas_cumprod)
    assert np.array_equal(ts, s.ddim_timesteps) and np.array_equal(a_prev, s.ddim_alphas_prev)
    with pytest.raises(ValueError):
        s.make_schedule(50, ddim_eta=0.5)


def test_grounding_input_null():
    from grounding_input.text_layout_tokinzer_input import GroundingNetInput
    gi = GroundingNetInput()
    with pytest.raises(AssertionError):
        gi.get_null_input()
    b = dict(boxes=torch.rand(2, 30, 4), masks=torch.ones(2, 30), text_embeddings=torch.randn(2, 30, 768))
    out = gi.prepare(b, None)
    assert out["positive_embeddings"] is b["text_embeddings"] and gi.max_box == 30
    z = gi.get_null_input()
    assert z["boxes"].shape == (2, 30, 4) and float(z["positive_embeddings"].abs().sum()) == 0.0


def test_ddim_name_importable_but_unusable():
    from ldm.models.diffusion.ddim import DDIMSampler
    with pytest.raises(NotImplementedError):
        DDIMSampler(None, None).sample()


def test_unknown_names_fail_with_clear_message():
    import ldm.modules.attention as att
    with pytest.raises(AttributeError, match="hot path"):
        att.ThisDoesNotExist


@pytest.mark.skipif(not os.path.isdir("/root/reference/GLIGEN"), reason="reference checkout not present (GPU box)")
def test_fallthrough_to_reference_for_names_outside_the_hot_path():
    """With the reference's GLIGEN/ further down sys.path (as txt2img.py arranges), names we do not own resolve there."""
    sys.path.append("/root/reference/GLIGEN")
    try:
        import ldm.modules.attention as att
        assert att.LinearAttention.__module__.startswith("_ltt_shadowed_")
        assert att.GatedSelfAttentionDense.__module__ == "ldm.modules.attention"
        import ldm.modules.diffusionmodules.util as u
        assert callable(u.checkpoint)                                   # reference-only helper
        from ldm.modules.distributions.distributions import DiagonalGaussianDistribution  # noqa: F401  (namespace merge)
    finally:
        sys.path.remove("/root/reference/GLIGEN")
