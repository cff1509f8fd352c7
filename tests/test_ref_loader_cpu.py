"""CPU checks of the staged reference (oracle/_ref): the files are byte-identical to /root/reference where that exists,
the reference tree and the drop-in tree coexist in one process, and the oracle port agrees with the reference modules
(fp32, tiny UNet) -- the same pin as tests/golden, but against the live reference code."""
import filecmp
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import make_ref, ref_loader as rl  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

needs_ref = pytest.mark.skipif(not rl.available(), reason="oracle/_ref not staged (no /root/reference in this container)")

TINY = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
            num_res_blocks=1, channel_mult=[1, 2], num_heads=8, transformer_depth=1, context_dim=768,
            fuser_type="gatedSA", grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)


@needs_ref
def test_staged_files_are_unmodified():
    if not os.path.isdir(make_ref.SRC):
        pytest.skip("/root/reference not present")
    for rel in make_ref.FILES:
        assert filecmp.cmp(os.path.join(make_ref.SRC, rel), os.path.join(make_ref.DST, rel), shallow=False), rel


@needs_ref
def test_reference_tree_coexists_with_dropin_and_matches_port():
    dropin = os.path.join(ROOT, "layoutllm_t2i_b200", "dropin")
    if dropin not in sys.path:
        sys.path.insert(0, dropin)
    import ldm.modules.attention as ours
    sd = uo.synthetic_state_dict(TINY, seed=7)
    ref = rl.build_unet(TINY, sd)
    import ldm.modules.attention as again
    assert again is ours and "dropin" in again.__file__          # the drop-in modules are back after the block
    assert os.path.join("oracle", "_ref") in rl._REF_MODULES[type(ref).__module__].__file__   # and `ref` is the reference class
    syn = uo.synthetic_inputs(B=2, H=16, W=16, n_boxes=3, seed=4321)
    syn["grounding"]["boxes"][1, 1] = torch.tensor([0.30, 0.2, 0.3001, 0.9])      # degenerate box: the `break` rule
    for cond in (True, False):
        for scale in (1.0, 0.35, 0.0):
            r = rl.unet_eps(ref, syn, 981, scale, cond, False)
            inp = dict(x=syn["x"], timesteps=torch.full((2,), 981), relations=syn["relations"],
                       context=syn["context"] if cond else syn["uc"])
            if cond:
                inp["grounding_input"] = syn["grounding"]
            with torch.no_grad():
                o = uo.unet_forward(sd, TINY, inp, scale=scale)
            assert ((o - r).norm() / r.norm()).item() < 1e-5


@needs_ref
def test_reference_sampler_records_102_evaluations():
    """PLMSSampler of the reference on a stub model: 51 [cond, uncond] pairs, gate 1 for 15 steps (+ the Euler
    predictor's second evaluation), then 0."""
    class Stub(torch.nn.Module):
        first_conv_restorable = False

        def __init__(self):
            super().__init__()
            with rl.reference_tree():
                from ldm.modules.attention import GatedSelfAttentionDense
            self.g = GatedSelfAttentionDense(64, 768, 8, 8)

        def restore_first_conv_from_SD(self):
            pass

        def forward(self, inp):
            return 0.1 * inp["x"]
    stub = Stub()
    rec = rl.Recorder(stub)
    sampler = rl.build_sampler(rec, "cpu")
    x = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    inp = dict(x=x.clone(), timesteps=None, context=torch.zeros(1, 77, 768), relations=torch.zeros(1, 10, 768),
               grounding_input=dict(), inpainting_extra_input=None, grounding_extra_input=None)
    sampler.sample(S=50, shape=(1, 4, 8, 8), input=inp, uc=torch.zeros(1, 77, 768), guidance_scale=7.5)
    assert len(rec.calls) == 102
    assert sum(1 for c in rec.calls if c["scale"] == 1.0) == 32 and rec.calls[0]["t"] == 981 and rec.calls[-1]["t"] == 1


@needs_ref
def test_dropin_autoencoder_has_the_reference_state_dict_grammar():
    """`autoencoder.load_state_dict(saved_ckpt["autoencoder"])` is strict in the callers (txt2img.py:107)."""
    import contextlib
    import io
    dropin = os.path.join(ROOT, "layoutllm_t2i_b200", "dropin")
    if dropin not in sys.path:
        sys.path.insert(0, dropin)
    from ldm.models.autoencoder import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    with rl.reference_tree(), contextlib.redirect_stdout(io.StringIO()):
        from ldm.models.autoencoder import AutoencoderKL as Ref
        ref = Ref(dd, 4, 0.18215)
    ours = AutoencoderKL(dd, 4, 0.18215)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs) == list(os_) and all(rs[k].shape == os_[k].shape for k in rs)
    ours.load_state_dict(rs)          # strict
    with pytest.raises(RuntimeError, match="CUDA device only"):
        ours.decode(torch.zeros(1, 4, 8, 8))
