#!/usr/bin/env python
"""GPU: device time of the sm_100a VAE decode (+ fused uint8 post-processing and one pinned D2H) per image batch, next to
the reference AutoencoderKL.decode under fp16 autocast + the callers' per-sample conversion loop (txt2img.py:317-324)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vae_checks as vc  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B, h in ((1, 64), (8, 64), (4, 96)):
    z = vc.latents(B, h, h)
    dec = vc.ours(vc.SD_VAE, 2, False)
    ref = vc.reference_vae(vc.SD_VAE, 2)
    host = torch.empty(B, 8 * h, 8 * h, 3, dtype=torch.uint8).pin_memory()
    l0 = dec.launch_count
    dec.decode(z)
    nl = dec.launch_count - l0
    t_dec = timed(lambda: dec.decode(z))
    t_u8 = timed(lambda: dec.decode_to_uint8(z, host))

    def ref_path():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            s = ref.decode(z)
        return [(torch.clamp(x, min=-1, max=1) * 0.5 + 0.5).cpu().numpy().transpose(1, 2, 0) * 255 for x in s]
    t_ref = timed(ref_path, reps=3)
    print(f"B={B} latent {h}x{h}: decode {t_dec:.2f} ms ({t_dec / B:.2f} ms/image, {nl} launches, {2514.5 * (h / 64) ** 2 * B / t_dec:.0f} GFLOP/ms "
          f"= {2.5145 * (h / 64) ** 2 * B / t_dec * 1e3:.0f} TFLOP/s); decode + uint8 + pinned D2H {t_u8:.2f} ms; "
          f"reference decode (fp16 autocast) + per-sample conversion {t_ref:.2f} ms", flush=True)
