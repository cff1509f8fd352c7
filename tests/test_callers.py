"""The drop-in boundary, proven by running the reference's callers UNMODIFIED.

`txt2img.py` and `GLIGEN/interface.py` are imported from their byte-identical copies under oracle/_ref (staged by
oracle/make_ref.py) with the drop-in tree installed the way INTEGRATION.md describes (PYTHONPATH + import hook); the
packages missing offline (`omegaconf`, `sng_parser`, `clip`, `backoff`, `pytorch_lightning`) and the out-of-scope side
models (CLIP, VAE) are small deterministic stand-ins under tests/stubs.  A synthetic checkpoint in the reference's
checkpoint grammar goes through `load_ckpt` (txt2img.py:96-114), then `generate_one_image` (txt2img.py:258-324) and
`run_batch_images` (interface.py:479-548, B = 3, per-sample boxes) run on the GPU and the latents they hand to
`autoencoder.decode` must equal what a direct `Engine.plms_sample` call produces for the same inputs.
"""
import importlib
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(ROOT, "layoutllm_t2i_b200", "dropin")
STUBS = os.path.join(ROOT, "tests", "stubs")

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "txt2img.py")),
                               reason="oracle/_ref not staged (python oracle/make_ref.py in the build container)")

# one level, 320 channels (restore_first_conv_from_SD hard-codes a 4->320 conv, openaimodel.py:400), 64x64 latent
# (generate_one_image hard-codes randn(B,4,64,64), txt2img.py:264)
SMALL = dict(image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[1],
             num_res_blocks=1, channel_mult=[1], num_heads=8, transformer_depth=1, context_dim=768,
             fuser_type="gatedSA", grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)


def _install():
    for p in (STUBS, DROPIN):
        if p not in sys.path:
            sys.path.insert(0, p)
    if REF not in sys.path:
        sys.path.append(REF)
    import _ltt_dropin_hook
    _ltt_dropin_hook.install()


def _import_callers():
    """import txt2img / interface exactly as `python txt2img.py` / `from GLIGEN.interface import ...` would."""
    _install()
    cwd = os.getcwd()
    os.chdir(REF)               # txt2img.py appends the RELATIVE path "./GLIGEN" (txt2img.py:15)
    try:
        txt2img = importlib.import_module("txt2img")
        interface = importlib.import_module("GLIGEN.interface")
    finally:
        os.chdir(cwd)
    gl = os.path.join(REF, "GLIGEN")
    if gl not in sys.path:
        sys.path.append(gl)      # keep "./GLIGEN" resolvable after the chdir back (SD_input_conv_weight_bias.pth lookup)
    return txt2img, interface


def _checkpoint(tmp_path):
    from oracle import unet_oracle as uo
    from ldm.models.diffusion.ldm import LatentDiffusion
    import ltt_test_stubs as st
    cfg = dict(
        model=dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
                   params=dict(image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[1],
                               num_res_blocks=1, channel_mult=[1], num_heads=8, transformer_depth=1, context_dim=768,
                               fuser_type="gatedSA", use_checkpoint=True,
                               grounding_tokenizer=dict(target="ldm.modules.diffusionmodules.text_grounding_net.PositionNet",
                                                        params=dict(in_dim=768, out_dim=768)))),
        autoencoder=dict(target="ltt_test_stubs.Autoencoder", params={}),
        text_encoder=dict(target="ltt_test_stubs.TextEncoder", params={}),
        diffusion=dict(target="ldm.models.diffusion.ldm.LatentDiffusion",
                       params=dict(linear_start=0.00085, linear_end=0.012, timesteps=1000)),
        grounding_tokenizer_input=dict(target="grounding_input.text_layout_tokinzer_input.GroundingNetInput"),
        max_relations=10)
    sd = uo.synthetic_state_dict(SMALL, seed=3)
    ckpt = dict(config_dict=dict(_content=cfg), model=sd, autoencoder=st.Autoencoder().state_dict(),
                text_encoder=st.TextEncoder().state_dict(),
                diffusion=LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).state_dict())
    path = os.path.join(str(tmp_path), "synthetic_ckpt.pth")
    torch.save(ckpt, path)
    return path, sd


@needs_ref
def test_unmodified_callers_import_the_dropin_tree(tmp_path):
    """CPU: both callers import, `load_ckpt` instantiates the drop-in classes from the checkpoint's config strings and
    loads the reference-grammar state_dict strictly by name (no GPU work)."""
    txt2img, interface = _import_callers()
    for mod in (txt2img, interface):
        assert os.path.join("oracle", "_ref") in mod.__file__
        assert DROPIN in sys.modules[mod.PLMSSampler.__module__].__file__, mod.PLMSSampler
        assert DROPIN in sys.modules[mod.instantiate_from_config.__module__].__file__
    path, sd = _checkpoint(tmp_path)
    for loader in (txt2img.load_ckpt, interface.load_ckpt):
        model, autoencoder, text_encoder, diffusion, config = loader(path, device="cpu")
        assert DROPIN in sys.modules[type(model).__module__].__file__
        assert DROPIN in sys.modules[type(diffusion).__module__].__file__
        got = model.state_dict()
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
        assert text_encoder.device == "cpu"          # the `'device' in vars(m)` probe of txt2img.py:112-114 fired
    # set_alpha_scale of the callers finds the drop-in fuser type by exact class (txt2img.py:46-50)
    txt2img.set_alpha_scale(model, 0.25)
    assert model.fuser_scale() == 0.25


def _direct(sd, rec, guidance=7.5):
    """The same sampling through the C-ABI wrapper only (no drop-in modules): Engine.plms_sample."""
    from layoutllm_t2i_b200.engine import Engine
    from oracle import plms_oracle as po
    e = Engine(SMALL, 0)
    e.load_state_dict(sd)
    e.finalize()
    x, ctx, rel, g, uc = rec["x"], rec["context"], rec["relations"], rec["grounding"], rec["uc"]
    B = x.shape[0]
    e.set_conditioning(torch.cat([ctx, uc]), torch.cat([rel, rel]), g, 64, 64, B)
    ts, a_t, a_prev, s1m = po.plms_tables(50, po.alphas_cumprod())
    w = torch.load(os.path.join(REF, "GLIGEN", "SD_input_conv_weight_bias.pth"), map_location="cpu")
    out = e.plms_sample(x, ts, a_t, a_prev, s1m, po.alpha_schedule(50), guidance, (w["weight"], w["bias"]))
    e.close()
    return out


@needs_ref
@pytest.mark.gpu
def test_generate_one_image_and_run_batch_images_unmodified(tmp_path, monkeypatch):
    txt2img, interface = _import_callers()
    import ltt_test_stubs as st
    from ldm.models.diffusion import plms as dropin_plms
    path, sd = _checkpoint(tmp_path)
    dev = torch.device("cuda", 0)
    recs = []
    orig = dropin_plms.PLMSSampler.sample

    def spy(self, S, shape, input, uc=None, guidance_scale=1, mask=None, x0=None):
        g = input["grounding_input"]
        recs.append(dict(x=input["x"].detach().clone(), context=input["context"].detach().clone(),
                         relations=input["relations"].detach().clone(), uc=uc.detach().clone(),
                         grounding={k: v.detach().clone() for k, v in g.items()}, S=S, guidance=guidance_scale))
        return orig(self, S, shape, input, uc, guidance_scale, mask, x0)
    monkeypatch.setattr(dropin_plms.PLMSSampler, "sample", spy)

    # ---- txt2img.generate_one_image (B = 1)
    gligen = list(txt2img.load_all_models(path, dev))
    cfg = gligen[4]
    cfg.update(dict(batch_size=1, no_plms=False, guidance_scale=7.5))
    gligen[4] = txt2img.OmegaConf.create(cfg)
    args = SimpleNamespace(batch_size=1)
    torch.manual_seed(11)
    imgs = txt2img.generate_one_image(args, tuple(gligen), "a cat on a sofa near a lamp", ["cat", "sofa", "lamp"],
                                      [[0.1, 0.3, 0.5, 0.8], [0.0, 0.5, 1.0, 1.0], [0.7, 0.1, 0.95, 0.6]],
                                      clip_model=st.ClipModel(), clip_processor=st.ClipProcessor(), device=dev)
    assert len(imgs) == 1 and imgs[0].size == (512, 512)
    z1 = gligen[1].decoded[-1]
    assert recs[-1]["S"] == 50 and z1.shape == (1, 4, 64, 64) and torch.isfinite(z1).all()
    assert recs[-1]["relations"].abs().sum() > 0           # the relation triplets reached the sampler
    assert torch.equal(z1.float().cpu(), _direct(sd, recs[-1]).cpu())
    model = gligen[0]
    assert model.fuser_scale() == 0 and model._sd_conv_active       # module state left as the reference loop leaves it

    # ---- GLIGEN/interface.run_batch_images (train_rl.py path: B = 3, per-sample prompts and boxes)
    all_models = interface.load_all_models(path, dev)
    meta = dict(prompts=["a dog under a table", "two birds", "a cup on a desk beside a book"],
                phrases=[["dog", "table"], ["bird"], ["cup", "desk", "book"]],
                locations=[[[0.2, 0.5, 0.6, 0.9], [0.1, 0.2, 0.9, 0.95]], [[0.3, 0.3, 0.7, 0.7]],
                           [[0.4, 0.4, 0.6, 0.7], [0.0, 0.6, 1.0, 1.0], [0.65, 0.35, 0.9, 0.65]]],
                alpha_type=[0.3, 0.0, 0.7])
    noise = torch.randn(3, 4, 64, 64, generator=torch.Generator().manual_seed(5)).to(dev)
    imgs = interface.run_batch_images(all_models, dict(batch_size=3, no_plms=False, guidance_scale=7.5), meta,
                                      noise.clone(), st.ClipModel(), st.ClipProcessor(), device=dev)
    assert len(imgs) == 3
    z3 = all_models[1].decoded[-1]
    rec = recs[-1]
    assert rec["x"].shape[0] == 3 and float(rec["grounding"]["masks"].sum()) == 6.0
    assert torch.equal(z3.float().cpu(), _direct(sd, rec).cpu())
    # per-sample conditioning matters: sample 1 alone (different prompt/boxes than its neighbours) reproduces row 1
    solo = {k: (v[1:2] if torch.is_tensor(v) else v) for k, v in rec.items() if k != "grounding"}
    solo["grounding"] = {k: v[1:2] for k, v in rec["grounding"].items()}
    z_solo = _direct(sd, solo)
    err = ((z_solo.cpu() - z3[1:2].float().cpu()).norm() / z3[1:2].float().norm().cpu()).item()
    assert err < 2e-2, err          # same arithmetic, other batch geometry (tile shapes): fp16 noise over 50 steps only


# ------------------------------------------------------------------------------------------------------------------
# every model of the checkpoint from the drop-in tree: UNet + sampler (rows a / b), AutoencoderKL (f1 / f2) and the CLIP text
# tower (f3) -- text encoder instantiated by load_ckpt from its reference config string, the callers' CLIPModel argument
# replaced by ClipModelAdapter
CLIP_SMALL = dict(vocab_size=1000, hidden_size=768, num_attention_heads=12, num_hidden_layers=2, intermediate_size=1024)
VAE_SMALL = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64, ch_mult=[1, 2],
                 num_res_blocks=1, attn_resolutions=[], dropout=0.0)


def _full_checkpoint(tmp_path):
    from oracle import clip_text_oracle as co
    from ldm.models.autoencoder import AutoencoderKL
    path, sd = _checkpoint(tmp_path)
    ck = torch.load(path, weights_only=False)
    cfg = ck["config_dict"]["_content"]
    cfg["text_encoder"] = dict(target="ldm.modules.encoders.modules.FrozenCLIPEmbedder", params=dict(text_config=CLIP_SMALL))
    cfg["autoencoder"] = dict(target="ldm.models.autoencoder.AutoencoderKL",
                              params=dict(ddconfig=VAE_SMALL, embed_dim=4, scale_factor=0.18215))
    ccfg = dict(co.default_clip_text_config(), **CLIP_SMALL)
    clip_sd = co.random_state_dict(ccfg, seed=8, with_projection=False)
    ck["text_encoder"] = {"transformer." + k: v for k, v in clip_sd.items()}
    torch.manual_seed(4)
    ck["autoencoder"] = AutoencoderKL(VAE_SMALL, 4, 0.18215).state_dict()
    torch.save(ck, path)
    return path, sd, ccfg, clip_sd


@needs_ref
@pytest.mark.gpu
def test_full_stack_generate_one_image_unmodified(tmp_path, monkeypatch):
    """Unmodified txt2img.generate_one_image with NOTHING stubbed on the device: drop-in UNet + PLMSSampler, drop-in
    AutoencoderKL, drop-in FrozenCLIPEmbedder, ClipModelAdapter as `clip_model`.  The conditioning the reference's
    per-string call sequence hands to the sampler must match ONE batched prepare_conditioning pass (2e-3: tile shapes differ),
    and both must match the fp32 oracle of the text tower."""
    txt2img, _ = _import_callers()
    import ltt_test_stubs as st
    import sng_parser
    from ldm.models.diffusion import plms as dropin_plms
    from layoutllm_t2i_b200.clip import ClipModelAdapter, prepare_conditioning, relation_phrases
    from oracle import clip_text_oracle as co
    from oracle.ref_loader import true_fp32
    path, sd, ccfg, clip_sd = _full_checkpoint(tmp_path)
    dev = torch.device("cuda", 0)
    recs = []
    orig = dropin_plms.PLMSSampler.sample

    def spy(self, S, shape, input, uc=None, guidance_scale=1, mask=None, x0=None):
        recs.append(dict(context=input["context"].detach().clone(), relations=input["relations"].detach().clone(),
                         uc=uc.detach().clone(), grounding={k: v.detach().clone() for k, v in input["grounding_input"].items()}))
        return orig(self, S, shape, input, uc, guidance_scale, mask, x0)
    monkeypatch.setattr(dropin_plms.PLMSSampler, "sample", spy)

    gligen = list(txt2img.load_all_models(path, dev))
    model, autoencoder, text_encoder = gligen[0], gligen[1], gligen[2]
    for m in (model, autoencoder, text_encoder):
        assert DROPIN in sys.modules[type(m).__module__].__file__, type(m)
    assert text_encoder.device == dev                  # txt2img.py:112-114
    text_encoder.set_tokenizer(st.HashTokenizer(ccfg["vocab_size"]))
    cfg = gligen[4]
    cfg.update(dict(batch_size=1, no_plms=False, guidance_scale=7.5))
    gligen[4] = txt2img.OmegaConf.create(cfg)
    prompt, phrases = "a cat on a sofa near a lamp", ["cat", "sofa", "lamp"]
    boxes = [[0.1, 0.3, 0.5, 0.8], [0.0, 0.5, 1.0, 1.0], [0.7, 0.1, 0.95, 0.6]]
    torch.manual_seed(11)
    imgs = txt2img.generate_one_image(SimpleNamespace(batch_size=1), tuple(gligen), prompt, phrases, boxes,
                                      clip_model=ClipModelAdapter(text_encoder.engine()), clip_processor=st.HashProcessor(ccfg["vocab_size"]),
                                      device=dev)
    assert len(imgs) == 1 and imgs[0].size == (128, 128)          # 2-level test VAE: 64 -> 128
    rec = recs[-1]

    # ---- ONE batched pass
    rels = relation_phrases(sng_parser.parse(prompt), cfg["max_relations"])
    assert len(rels) == 5                               # PAD + 2 triplets twice
    out = prepare_conditioning(text_encoder.engine(), text_encoder.tokenize_padded, prompt, phrases, boxes, rels, batch=1,
                               max_relas=cfg["max_relations"])

    def rel_err(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())
    assert rel_err(out["context"], rec["context"]) < 2e-3 and rel_err(out["uc"], rec["uc"]) < 2e-3
    assert rel_err(out["relations"], rec["relations"]) < 2e-3 and float(rec["relations"][0, 5:].abs().max()) == 0.0
    assert rel_err(out["text_embeddings"], rec["grounding"]["positive_embeddings"]) < 2e-3
    assert torch.equal(out["boxes"], rec["grounding"]["boxes"]) and torch.equal(out["masks"], rec["grounding"]["masks"])

    # ---- both against the fp32 oracle of the text tower
    sdd = {k: v.to(dev) for k, v in clip_sd.items()}
    with torch.no_grad(), true_fp32():
        z, _ = co.clip_text_forward(sdd, ccfg, text_encoder.tokenize_padded([prompt, ""]).to(dev))
        pooled = torch.cat([co.clip_text_forward(sdd, ccfg, st.HashTokenizer(ccfg["vocab_size"])(p, padding=True)["input_ids"].to(dev))[1]
                            for p in phrases])
    assert rel_err(rec["context"][0], z[0]) < 2e-3 and rel_err(rec["uc"][0], z[1]) < 2e-3
    assert rel_err(rec["grounding"]["positive_embeddings"][0, :3], pooled) < 2e-3
