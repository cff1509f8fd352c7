#!/usr/bin/env python
"""GPU: parity of the sm_100a engine against the UNMODIFIED reference under CUDA fp16 autocast (oracle/_ref).

Writes gpurun_out/parity_r02.json + .txt:
  * the reference's own fp16 noise floor (same sample, settings that leave its arithmetic specification unchanged),
  * engine vs reference at BASELINE configs 1-4,
  * the teacher-forced 50-step PLMS table: the reference sampler drives, every recorded x_t is re-evaluated
    by the engine; per-evaluation rel-L2 of eps (cond, uncond, CFG-combined) + the free-running final latent.

    python tools/gpu_parity_steps.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import ref_checks as rc  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    out = dict(device=torch.cuda.get_device_name(0), torch=torch.__version__)
    FULL = rc.FULL
    out["reference_noise_floor_64x64_gate1"] = rc.reference_noise(FULL, 0, 64, 64, 6, 481, 1.0)
    out["reference_noise_floor_64x64_gate0"] = rc.reference_noise(FULL, 0, 64, 64, 6, 481, 0.0)
    cases = [("config1 64x64 B=1 2 boxes t=981 gate1", dict(B=1, H=64, W=64, n_boxes=2, t=981, scale=1.0)),
             ("config2 64x64 B=1 6 boxes t=481 gate1", dict(B=1, H=64, W=64, n_boxes=6, t=481, scale=1.0)),
             ("config2 64x64 B=1 6 boxes t=481 gate0", dict(B=1, H=64, W=64, n_boxes=6, t=481, scale=0.0))]
    if not quick:
        cases += [("config3 64x64 B=8 1 box", dict(B=8, H=64, W=64, n_boxes=1, t=601, scale=1.0, distinct=True)),
                  ("config3 64x64 B=8 30 boxes", dict(B=8, H=64, W=64, n_boxes=30, t=601, scale=1.0, distinct=True)),
                  ("config4 96x96 B=4 6 boxes gate1", dict(B=4, H=96, W=96, n_boxes=6, t=801, scale=1.0)),
                  ("config4 96x96 B=4 6 boxes gate0", dict(B=4, H=96, W=96, n_boxes=6, t=201, scale=0.0))]
    out["engine_vs_reference_fp16"] = {}
    for name, kw in cases:
        c, u = rc.engine_vs_reference(cfg=FULL, seed=0, **kw)
        out["engine_vs_reference_fp16"][name] = dict(cond=c, uncond=u)
        print(name, c, u, flush=True)
    c, u = rc.engine_vs_reference(cfg=FULL, seed=0, B=1, H=64, W=64, n_boxes=2, t=981, scale=1.0, autocast=False)
    out["engine_vs_reference_fp32"] = dict(cond=c, uncond=u)
    rows, free, ref_final, ours = rc.teacher_forced_table(FULL, 0, 1, 64, 64, 6, 50)
    out["teacher_forced_50_steps"] = rows
    out["free_running_final_latent_rel_l2"] = free
    # the reference's own free-running noise: the same 50 steps with the sample batched with a second one is not available
    # through the sampler API, so the floor for the final latent is the fp32 run of the reference sampler
    syn, calls, ref32, _, _ = rc.reference_trace(FULL, 0, 1, 64, 64, 6, 50, autocast=False)
    out["free_running_reference_fp16_vs_fp32"] = rc.rel(ref_final, ref32)
    out["free_running_engine_vs_reference_fp32"] = rc.rel(ours, ref32)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_r02.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "parity_r02.txt"), "w") as f:
        f.write(f"{out['device']}, torch {out['torch']}; rel-L2 vs the unmodified reference under torch.autocast('cuda', fp16)\n")
        for k in ("reference_noise_floor_64x64_gate1", "reference_noise_floor_64x64_gate0"):
            f.write(f"\n{k} (reference vs ITSELF):\n")
            for n, v in out[k].items():
                f.write(f"  {n:48s} {v:.3e}\n")
        f.write("\nengine vs reference fp16 (cond / uncond):\n")
        for n, v in out["engine_vs_reference_fp16"].items():
            f.write(f"  {n:48s} {v['cond']:.3e} / {v['uncond']:.3e}\n")
        f.write(f"  engine vs reference fp32, config 1:              {c:.3e} / {u:.3e}\n")
        f.write("\nteacher-forced 50-step PLMS (reference sampler drives; engine re-evaluates each x_t):\n")
        f.write("  eval    t  gate  first_conv   eps_cond   eps_uncond  eps_cfg\n")
        for r in rows:
            f.write(f"  {r['eval']:4d} {r['t']:4d}  {r['gate']:.1f}   {r['first_conv']:7s}   {r['eps_cond']:.3e}  {r['eps_uncond']:.3e}  {r['eps_cfg']:.3e}\n")
        f.write(f"\nfree-running 50 steps, final latent: engine vs reference fp16 {free:.3e}; reference fp16 vs reference fp32 "
                f"{out['free_running_reference_fp16_vs_fp32']:.3e}; engine vs reference fp32 {out['free_running_engine_vs_reference_fp32']:.3e}\n")
    print(open(os.path.join(ROOT, "gpurun_out", "parity_r02.txt")).read())


if __name__ == "__main__":
    main()
