"""LatentDiffusion = schedule buffers + q_sample (reference GLIGEN/ldm/models/diffusion/ldm.py:11-24)."""
import torch

from ldm.modules.diffusionmodules.util import extract_into_tensor
from .ddpm import DDPM


class LatentDiffusion(DDPM):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.clip_denoised = False

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)
