"""Host-side helpers of the diffusion path (reference GLIGEN/ldm/modules/diffusionmodules/util.py): noise-schedule
tables and the layer factories the parameter containers use.  The embedders and norms themselves execute inside
the sm_100a library; names not defined here fall through to the reference's file."""
import math

import numpy as np
import torch
import torch.nn as nn

import _ltt_fallthrough


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """fp64 numpy betas (reference util.py:30-52)."""
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        steps = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        a = torch.cos(steps / (1 + cosine_s) * math.pi / 2) ** 2
        a = a / a[0]
        betas = (1 - a[1:] / a[:-1]).clamp(0, 0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """Sub-sampled step indices, shifted by one (reference util.py:55-70)."""
    if ddim_discr_method == "uniform":
        steps = np.arange(0, num_ddpm_timesteps, num_ddpm_timesteps // num_ddim_timesteps)
    elif ddim_discr_method == "quad":
        steps = (np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return steps + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """(sigmas, alphas, alphas_prev) for the sub-sampled steps (reference util.py:73-83): `alphas` keeps the dtype of
    `alphacums` (torch fp32), `alphas_prev` is a numpy fp64 array built from python floats."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def extract_into_tensor(a, t, x_shape):
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


class GroupNorm32(nn.GroupNorm):
    """32-group GroupNorm whose statistics are taken in fp32 (reference util.py:226-229).  Parameter container: the
    normalisation runs inside the library's fused kernels."""


def normalization(channels):
    return GroupNorm32(32, channels)


def conv_nd(dims, *args, **kwargs):
    if dims != 2:
        raise ValueError("the B200 path implements the 2-D UNet only")
    return nn.Conv2d(*args, **kwargs)


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


_ltt_fallthrough.install(__name__, globals())
