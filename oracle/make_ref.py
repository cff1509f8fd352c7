"""Recipe: stage the UNMODIFIED reference files of the hot path under the git-ignored ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY (see oracle/unet_oracle.py).  The reference is pure Python, so "building" it means copying
the files where they lie under /root/reference into a directory that travels to the GPU box with the repository
snapshot (``oracle/_ref/`` is listed in .gitignore, not in .gpurunignore: it never enters history, it does ship).
Nothing is edited: the copies are byte-identical (checked below), so a test that runs them runs the reference.

Used for
  * parity pinning on the GPU: the reference UNetModel / PLMSSampler / AutoencoderKL under ``torch.autocast('cuda')``
    is the "reference fp16 output" north_star names (tests/ref_checks.py, tools/gpu_parity_steps.py);
  * the boundary proof: ``txt2img.py`` and ``GLIGEN/interface.py`` are imported unmodified with the drop-in tree ahead
    of ``oracle/_ref/GLIGEN`` on sys.path (tests/test_callers_gpu.py);
  * ``bench.py --impl reference`` (kind "reference": the reference's own modules on the host cores).

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference exists
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LTT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

# paths relative to the reference root
FILES = [
    # callers that must run unchanged (boundary)
    "txt2img.py", "utils.py", "base_prompt.py", "models/policy.py", "models/llm.py", "tools/aesthetic.py", "tools/metrics.py",
    "GLIGEN/interface.py", "GLIGEN/trainer.py", "GLIGEN/inpaint_mask_func.py", "GLIGEN/distributed.py",
    "GLIGEN/dataset/__init__.py", "GLIGEN/dataset/concat_dataset.py", "GLIGEN/dataset/catalog.py",
    "GLIGEN/SD_input_conv_weight_bias.pth",
    # the hot path (a1-a18)
    "GLIGEN/ldm/util.py",
    "GLIGEN/ldm/modules/attention.py",
    "GLIGEN/ldm/modules/diffusionmodules/__init__.py",
    "GLIGEN/ldm/modules/diffusionmodules/openaimodel.py",
    "GLIGEN/ldm/modules/diffusionmodules/util.py",
    "GLIGEN/ldm/modules/diffusionmodules/text_grounding_net.py",
    "GLIGEN/ldm/models/diffusion/__init__.py",
    "GLIGEN/ldm/models/diffusion/plms.py",
    "GLIGEN/ldm/models/diffusion/ddim.py",
    "GLIGEN/ldm/models/diffusion/ddpm.py",
    "GLIGEN/ldm/models/diffusion/ldm.py",
    "GLIGEN/grounding_input/__init__.py",
    "GLIGEN/grounding_input/text_layout_tokinzer_input.py",
    # next row f1: the VAE decoder
    "GLIGEN/ldm/models/autoencoder.py",
    "GLIGEN/ldm/modules/diffusionmodules/model.py",
    "GLIGEN/ldm/modules/distributions/__init__.py",
    "GLIGEN/ldm/modules/distributions/distributions.py",
    # next row f3: the text encoder wrapper (the transformer itself is the third-party `transformers` package)
    "GLIGEN/ldm/modules/encoders/__init__.py",
    "GLIGEN/ldm/modules/encoders/modules.py",
    "GLIGEN/ldm/modules/x_transformer.py",
]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box?) -- keeping the staged copy under {DST}")
        return 0
    n = 0
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            print(f"make_ref: missing in reference: {rel}", file=sys.stderr)
            return 1
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.isfile(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
            n += 1
        assert filecmp.cmp(s, d, shallow=False), rel
    print(f"make_ref: {len(FILES)} reference files staged under {DST} ({n} refreshed), byte-identical")
    return 0


if __name__ == "__main__":
    sys.exit(main())
