"""Run the UNMODIFIED reference modules staged under oracle/_ref (see oracle/make_ref.py) next to the drop-in tree.

TEST INFRASTRUCTURE ONLY.  Both trees define the top-level namespace packages ``ldm`` and ``grounding_input``; this
module imports the reference's copies with ``sys.modules`` / ``sys.path`` swapped for the duration of a ``with
reference_tree():`` block, so that the reference's own absolute imports (``from ldm.modules... import``) resolve inside
the reference tree and the drop-in modules already imported by the process are put back afterwards.  Objects built
inside the block keep working outside it (their classes hold their own module globals).

Everything the reference builds from strings (``instantiate_from_config``) must be constructed inside the block.
"""
from __future__ import annotations

import contextlib
import os
import sys
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
REF_GLIGEN = os.path.join(REF_ROOT, "GLIGEN")
_PKGS = ("ldm", "grounding_input", "_ltt_fallthrough")
_REF_MODULES: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REF_GLIGEN, "ldm", "modules", "diffusionmodules", "openaimodel.py"))


def _ours(path: str) -> bool:
    return os.path.abspath(path or ".").rstrip("/").endswith(os.path.join("layoutllm_t2i_b200", "dropin"))


@contextlib.contextmanager
def reference_tree():
    if not available():
        raise FileNotFoundError(f"{REF_GLIGEN} is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _PKGS}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_REF_MODULES)
    saved_path = list(sys.path)
    sys.path[:] = [REF_GLIGEN] + [p for p in saved_path if not _ours(p)]
    # the drop-in import hook (layoutllm_t2i_b200/dropin/_ltt_dropin_hook.py) would serve its own modules: park it
    hooks = [f for f in sys.meta_path if type(f).__name__ == "DropinFinder"]
    for f in hooks:
        sys.meta_path.remove(f)
    try:
        yield
    finally:
        sys.meta_path[:0] = hooks
        for k in [k for k in sys.modules if k.split(".")[0] in _PKGS]:
            _REF_MODULES[k] = sys.modules.pop(k)
        sys.modules.update(saved)
        sys.path[:] = saved_path


@contextlib.contextmanager
def true_fp32():
    """fp32 reference runs must be fp32: PyTorch lets cuDNN convolutions use TF32 on CUDA by default (10-bit mantissa,
    7.8e-4 rel-L2 on this UNet), which is not the arithmetic of the reference's CPU / fp32 path."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m


def unet_params(cfg: dict) -> dict:
    """Constructor arguments of the reference UNetModel for an oracle-style cfg dict (GLIGEN/configs/coco2014.yaml:8-30)."""
    return dict(image_size=cfg.get("image_size", 64), in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                model_channels=cfg["model_channels"], attention_resolutions=list(cfg["attention_resolutions"]),
                num_res_blocks=cfg["num_res_blocks"], channel_mult=list(cfg["channel_mult"]), num_heads=cfg["num_heads"],
                transformer_depth=1, context_dim=cfg["context_dim"], fuser_type="gatedSA", use_checkpoint=False,
                grounding_tokenizer=dict(target="ldm.modules.diffusionmodules.text_grounding_net.PositionNet",
                                         params=dict(in_dim=cfg.get("grounding_in_dim", 768), out_dim=cfg.get("grounding_out_dim", 768))))


def build_unet(cfg: dict, sd: dict, device="cpu"):
    """The reference's UNetModel (openaimodel.py:235-391) with `sd` loaded as txt2img.py:106 does."""
    with reference_tree():
        from grounding_input.text_layout_tokinzer_input import GroundingNetInput
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        m = UNetModel(**unet_params(cfg)).eval()
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing[:4], unexpected[:4])
        m.grounding_tokenizer_input = GroundingNetInput()
    return m.to(device)


def set_alpha_scale(model, alpha_scale):
    """txt2img.py:46-50 on reference module types."""
    for module in model.modules():
        if type(module).__name__ in ("GatedCrossAttentionDense", "GatedSelfAttentionDense"):
            module.scale = alpha_scale


def alpha_generator(length, type=(0.3, 0.0, 0.7)):
    """txt2img.py:59-93."""
    import numpy as np
    n0, n1 = int(type[0] * length), int(type[1] * length)
    n2 = length - n0 - n1
    decay = list(np.arange(0, 1, 1 / n1)[::-1]) if n1 else []
    return [1] * n0 + decay + [0] * n2


def model_inputs(model, syn: dict, t: int, cond: bool) -> dict:
    """The dict UNetModel.forward takes (txt2img.py:302-310 / plms.py:118-121) from oracle-style synthetic inputs."""
    B = syn["x"].shape[0]
    inp = dict(x=syn["x"], timesteps=torch.full((B,), t, dtype=torch.long, device=syn["x"].device),
               context=syn["context"] if cond else syn["uc"], relations=syn["relations"],
               inpainting_extra_input=None, grounding_extra_input=None)
    if cond:
        g = syn["grounding"]
        inp["grounding_input"] = model.grounding_tokenizer_input.prepare(
            dict(boxes=g["boxes"], masks=g["masks"], text_embeddings=g["positive_embeddings"]), None)
    else:       # get_null_input() replays the sizes remembered by the last prepare() (text_layout_tokinzer_input.py:47-62)
        g = syn["grounding"]
        model.grounding_tokenizer_input.prepare(dict(boxes=g["boxes"], masks=g["masks"], text_embeddings=g["positive_embeddings"]), None)
    return inp


@torch.no_grad()
def unet_eps(model, syn: dict, t: int, scale: float, cond: bool, autocast: bool) -> torch.Tensor:
    set_alpha_scale(model, scale)
    inp = model_inputs(model, syn, t, cond)
    if autocast:
        with torch.autocast(syn["x"].device.type, dtype=torch.float16):
            return model(inp).float()
    with true_fp32():
        return model(inp).float()


def build_sampler(model, device, alpha_type=(0.3, 0.0, 0.7)):
    """PLMSSampler + LatentDiffusion of the reference, wired as txt2img.py:285-287."""
    with reference_tree():
        from ldm.models.diffusion.ldm import LatentDiffusion
        from ldm.models.diffusion.plms import PLMSSampler
        diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(device)
        return PLMSSampler(diffusion, model, alpha_generator_func=partial(alpha_generator, type=list(alpha_type)),
                           set_alpha_scale=set_alpha_scale)


class Recorder(torch.nn.Module):
    """Wraps the reference UNet inside the reference sampler and records every evaluation (teacher forcing)."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.calls = []

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("model"), name)

    def restore_first_conv_from_SD(self):
        return self.model.restore_first_conv_from_SD()

    def forward(self, inp):
        out = self.model(inp)
        scale = next(m.scale for m in self.model.modules() if type(m).__name__ == "GatedSelfAttentionDense")
        self.calls.append(dict(x=inp["x"].detach().clone(), t=int(inp["timesteps"][0]), cond="grounding_input" in inp,
                               scale=float(scale), eps=out.detach().float().clone(),
                               first_conv=getattr(self.model, "first_conv_type", "SD") if hasattr(self.model, "GLIGEN_first_conv_state_dict") else None))
        return out
