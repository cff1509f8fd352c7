"""CPU restatement of the reference's PLMS sampler (schedule + 50-step loop).

TEST INFRASTRUCTURE ONLY -- see the header of ``oracle/unet_oracle.py``.
Follows ``/root/reference/GLIGEN/ldm/models/diffusion/plms.py`` (cited per
function); pinned against the reference's own sampler by
``tests/gen_golden.py`` -> ``tests/golden/plms_*.pt``.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch


def beta_schedule_linear(n: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> np.ndarray:
    """make_beta_schedule("linear") in fp64 (ldm/modules/diffusionmodules/util.py:30-34)."""
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2).numpy()


def alphas_cumprod(n: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012) -> torch.Tensor:
    """DDPM.register_schedule: cumprod in fp64 then stored as fp32 (ddpm.py:19-36)."""
    return torch.tensor(np.cumprod(1.0 - beta_schedule_linear(n, linear_start, linear_end), axis=0), dtype=torch.float32)


def plms_tables(S: int, acp: torch.Tensor):
    """PLMSSampler.make_schedule with eta=0 (plms.py:25-56; util.py:55-83).
    Returns (timesteps[S] int, a_t[S], a_prev[S], sqrt(1-a_t)[S]) as numpy."""
    T = acp.shape[0]
    ts = np.asarray(list(range(0, T, T // S))) + 1
    a = acp.cpu()[ts]                                              # torch fp32
    a_prev = np.asarray([acp[0].item()] + acp.cpu()[ts[:-1]].tolist())   # numpy fp64 of fp32 values
    return ts, a.numpy(), a_prev, np.sqrt(1.0 - a.numpy())


def alpha_schedule(length: int, kind=(0.3, 0.0, 0.7)) -> List[float]:
    """alpha_generator (txt2img.py:59-93): [1]*n0 + linear decay + [0]*n2."""
    n0, n1 = int(kind[0] * length), int(kind[1] * length)
    n2 = length - n0 - n1
    decay = list(np.arange(0, 1, 1 / n1)[::-1]) if n1 else []
    return [1] * n0 + decay + [0] * n2


def plms_sample(model_eps: Callable[[torch.Tensor, torch.Tensor, bool, float, bool], torch.Tensor],
                x: torch.Tensor, S: int = 50, guidance: float = 7.5, acp: Optional[torch.Tensor] = None,
                alpha_kind=(0.3, 0.0, 0.7), trace: Optional[list] = None) -> torch.Tensor:
    """plms_sampling + p_sample_plms (plms.py:64-163).

    `model_eps(x, t, cond, scale, sd_first_conv)` returns eps for the cond
    (grounding present) or uncond (null grounding, context=uc) branch; `scale`
    is the gate value of the step and `sd_first_conv` says whether
    restore_first_conv_from_SD has happened (it is permanent once alpha hits 0,
    plms.py:86-87 / openaimodel.py:393-405).
    """
    acp = alphas_cumprod() if acp is None else acp
    ts, a_t, a_prev, s1m = plms_tables(S, acp)
    sched = alpha_schedule(S, alpha_kind)
    B = x.shape[0]
    order = np.flip(ts)
    old: list = []
    restored = False

    def eps_cfg(xx, t, scale):
        tt = torch.full((B,), int(t), device=x.device, dtype=torch.long)
        e_c = model_eps(xx, tt, True, scale, restored)
        if guidance != 1:
            e_u = model_eps(xx, tt, False, scale, restored)
            e_c = e_u + guidance * (e_c - e_u)
        return e_c

    def step_to_prev(xx, e, idx):
        at = torch.full((B, 1, 1, 1), float(a_t[idx]), device=x.device)
        ap = torch.full((B, 1, 1, 1), float(a_prev[idx]), device=x.device)
        sq = torch.full((B, 1, 1, 1), float(s1m[idx]), device=x.device)
        pred_x0 = (xx - sq * e) / at.sqrt()
        torch.randn_like(xx)          # sigma == 0 but the reference still draws (plms.py:138)
        return ap.sqrt() * pred_x0 + (1.0 - ap).sqrt() * e

    for i, t in enumerate(order):
        scale = sched[i]
        if scale == 0:
            restored = True
        idx = S - 1 - i
        t_next = order[min(i + 1, S - 1)]
        e = eps_cfg(x, t, scale)
        if len(old) == 0:
            e_next = eps_cfg(step_to_prev(x, e, idx), t_next, scale)
            ep = (e + e_next) / 2
        elif len(old) == 1:
            ep = (3 * e - old[-1]) / 2
        elif len(old) == 2:
            ep = (23 * e - 16 * old[-1] + 5 * old[-2]) / 12
        else:
            ep = (55 * e - 59 * old[-1] + 37 * old[-2] - 9 * old[-3]) / 24
        x = step_to_prev(x, ep, idx)
        old.append(e)
        if len(old) >= 4:
            old.pop(0)
        if trace is not None:
            trace.append(x.clone())
    return x
