#!/usr/bin/env python
"""Headline benchmark: 512x512 images/sec of the layout-conditioned sampler (50 PLMS steps, CFG 7.5, 6 boxes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one complete image batch through the reference-facing API of the drop-in tree, as the reference's
generate_one_image does it (txt2img.py:302-324): `PLMSSampler.sample(S=50, shape, input, uc, guidance_scale=7.5)` on the
drop-in `UNetModel` (102 UNet evaluations per image: 51 [cond ; uncond] pairs), then `AutoencoderKL.decode` of the latents
with the uint8 / HWC image conversion fused in (the metric is sampler + decode; the sampler-only figure is
reported beside it).  `value` times K steps with the conditioning tensors already in HBM and the images left in HBM;
`e2e` times K more steps whose inputs start in pinned HOST memory (H2D inside the step) and whose uint8 images are read
back to the host with one pinned copy.  Weights are random-init tensors of the LayoutLLM-T2I architecture (no checkpoints offline).
N > 1: every rank samples its own batch slice (no data-path collective) and the final latents are all-gathered once
per step over NCCL; weak scaling (per-GPU batch fixed) unless --global-batch is given.

`--impl reference`: the reference's own UNetModel (byte-identical copies under oracle/_ref; fp32 torch CPU, all
host threads) on a bounded sample of the same workload: each step = one [cond, uncond] UNet evaluation pair at the
workload's size; images/sec = 1 / (51 pairs x t_pair).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "layoutllm_t2i_b200", "dropin"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

UNET_CFG = dict(image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
                fuser_type="gatedSA", grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)
# algorithmic FLOPs (2*MAC) of one UNet evaluation at B=1, from FlopCounterMode on the reference module
GF_FWD = {64: (1147.69, 814.14), 96: (3249.49, 2158.99)}        # latent size -> (alpha=1, alpha=0 with the fuser elided)
GF_VAE_DECODE_64 = 2514.5                                       # AutoencoderKL.decode of one 64x64 latent, ~ (lat/64)^2
VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                    num_res_blocks=2, attn_resolutions=[], dropout=0.0)          # GLIGEN/configs/coco2014.yaml:33-52
GUIDANCE = 7.5
ALPHA_TYPE = (0.3, 0.0, 0.7)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), tflops_burst=float(d.get("bf16_tflops")),
                        hbm=float(d.get("hbm_gbs")), source="measured (MEASURED_PEAKS.json, sustained bf16)")
        except Exception:  # noqa: BLE001
            pass
    return dict(tflops=1590.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def alpha_generator(length, type=ALPHA_TYPE):
    """Gate schedule of the reference's callers (txt2img.py:59-93)."""
    n0, n1 = int(type[0] * length), int(type[1] * length)
    n2 = length - n0 - n1
    decay = list(np.arange(0, 1, 1 / n1)[::-1]) if n1 else []
    return [1] * n0 + decay + [0] * n2


def set_alpha_scale(model, alpha_scale):
    """As the reference's callers define it (txt2img.py:46-50)."""
    from ldm.modules.attention import GatedCrossAttentionDense, GatedSelfAttentionDense
    for module in model.modules():
        if type(module) == GatedCrossAttentionDense or type(module) == GatedSelfAttentionDense:
            module.scale = alpha_scale


def synthetic_host_inputs(B, H, W, n_boxes, rank, pin):
    """Seeded synthetic batch of SURVEY.md 8(d) on the host (pinned when a GPU is present)."""
    g = lambda s: torch.Generator().manual_seed(1234 + s + 1000 * rank)
    x = torch.randn(B, 4, H, W, generator=g(0))
    context = torch.randn(1, 77, 768, generator=g(1)).repeat(B, 1, 1)
    uc = torch.randn(1, 77, 768, generator=g(2)).repeat(B, 1, 1)
    relations = torch.zeros(B, 10, 768)
    relations[:, :3] = torch.randn(1, 3, 768, generator=g(3))
    emb = torch.zeros(B, 30, 768)
    emb[:, :n_boxes] = torch.randn(B, n_boxes, 768, generator=g(4))
    boxes = torch.zeros(B, 30, 4)
    u = torch.rand(B, n_boxes, 4, generator=g(5))
    x0, y0 = u[..., 0] * 0.5, u[..., 1] * 0.5
    bw, bh = 0.2 + 0.3 * u[..., 2], 0.2 + 0.3 * u[..., 3]
    boxes[:, :n_boxes] = torch.stack([x0, y0, (x0 + bw).clamp(max=1.0), (y0 + bh).clamp(max=1.0)], dim=-1)
    masks = torch.zeros(B, 30)
    masks[:, :n_boxes] = 1
    d = dict(x=x, context=context, uc=uc, relations=relations, boxes=boxes, masks=masks, text_embeddings=emb)
    return {k: (v.pin_memory() if pin else v) for k, v in d.items()}


def build_model(device):
    """Drop-in UNetModel with random-init weights created directly on the device (gates non-zero so the gated
    self-attention and relation-fusion paths are live; the reference initialises them to 0)."""
    from grounding_input.text_layout_tokinzer_input import GroundingNetInput
    from ldm.util import instantiate_from_config
    cfg = dict(target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
               params=dict(image_size=64, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                           num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
                           fuser_type="gatedSA", use_checkpoint=True,
                           grounding_tokenizer=dict(target="ldm.modules.diffusionmodules.text_grounding_net.PositionNet",
                                                    params=dict(in_dim=768, out_dim=768))))
    with torch.device("meta"):
        model = instantiate_from_config(cfg)
    model = model.to_empty(device=device).eval()
    gen = torch.Generator(device=device).manual_seed(0)
    flip = 0
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("alpha_attn") or name.endswith("alpha_dense"):
                p.fill_(0.5 if flip % 2 == 0 else -0.4)
                flip += 1
            elif p.dim() >= 2:
                bound = 1.0 / math.sqrt(p[0].numel())
                p.copy_((torch.rand(p.shape, generator=gen, device=device) * 2 - 1) * bound)
            elif "norm" in name and name.endswith("weight") or name.endswith("layers.0.weight") or name == "out.0.weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen, device=device))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=gen, device=device))
    model.grounding_tokenizer_input = GroundingNetInput()
    g2 = torch.Generator().manual_seed(5)       # stand-in for SD_input_conv_weight_bias.pth (ships with the reference only)
    model.set_sd_first_conv(0.2 * torch.randn(320, 4, 3, 3, generator=g2), 0.02 * torch.randn(320, generator=g2))
    return model


def build_vae(device):
    """Drop-in AutoencoderKL (SD VAE geometry) with random-init weights on the device."""
    from ldm.util import instantiate_from_config
    torch.manual_seed(1)
    vae = instantiate_from_config(dict(target="ldm.models.autoencoder.AutoencoderKL",
                                       params=dict(ddconfig=VAE_DDCONFIG, embed_dim=4, scale_factor=0.18215)))
    return vae.to(device).eval()


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w=float(np.median(power)),
                    samples=len(sm), reasons=sorted(reasons))


# ---------------------------------------------------------------------------------------------------- reference arm
def reference_kind():
    """"reference": the reference's own UNetModel from the byte-identical copies under oracle/_ref (oracle/make_ref.py);
    "port": the oracle restatement (only when the staged copy is missing)."""
    from oracle import ref_loader as rl
    return "reference" if rl.available() else "port"


class ReferenceUNet:
    """The reference algorithm on a device of choice: one [cond, uncond] evaluation pair = two separate B-sample forwards,
    exactly as the reference sampler issues them (plms.py:116-122)."""

    def __init__(self, sd_cpu, dev):
        from oracle import ref_loader as rl
        self.kind, self.dev = reference_kind(), dev
        if self.kind == "reference":
            self.model = rl.build_unet(UNET_CFG, sd_cpu, dev)
        else:
            self.sd = {k: v.to(dev) for k, v in sd_cpu.items()}

    def pair(self, inputs, scale, autocast):
        from oracle import ref_loader as rl
        from oracle import unet_oracle as uo
        inp = {k: v.to(self.dev) for k, v in inputs.items()}
        syn = dict(x=inp["x"], context=inp["context"], uc=inp["uc"], relations=inp["relations"],
                   grounding=dict(boxes=inp["boxes"], masks=inp["masks"], positive_embeddings=inp["text_embeddings"]))
        if self.kind == "reference":
            rl.unet_eps(self.model, syn, 981, scale, True, autocast)
            rl.unet_eps(self.model, syn, 981, scale, False, autocast)
            return
        B = inp["x"].shape[0]
        ts = torch.full((B,), 981, dtype=torch.long, device=self.dev)
        cond = dict(x=inp["x"], timesteps=ts, context=inp["context"], relations=inp["relations"], grounding_input=syn["grounding"])
        unc = dict(x=inp["x"], timesteps=ts, context=inp["uc"], relations=inp["relations"])
        with torch.no_grad(), torch.autocast(torch.device(self.dev).type, dtype=torch.float16, enabled=autocast):
            uo.unet_forward(self.sd, UNET_CFG, cond, scale=scale)
            uo.unet_forward(self.sd, UNET_CFG, unc, scale=scale)


def cpu_pair_seconds(ref, inputs, scale, n_threads):
    """One [cond, uncond] UNet evaluation pair of the reference on the host (fp32, all threads)."""
    torch.set_num_threads(n_threads)
    t0 = time.perf_counter()
    ref.pair(inputs, scale, autocast=False)
    return time.perf_counter() - t0


def gpu_eager_pair_ms(ref, inputs, scale, reps=3):
    """One [cond, uncond] evaluation pair of the reference's own modules as EAGER PyTorch on the GPU under fp16 autocast
    (the reference's 1-GPU eager path of north_star's ">= 6x" target).  A reported baseline only: nothing of the
    product path runs here."""
    def pair():
        ref.pair(inputs, scale, autocast=True)
    for _ in range(2):
        pair()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pair()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_state_dict():
    """Random-init fp32 weights of the architecture on the host (values do not affect CPU timing)."""
    from oracle import unet_oracle as uo
    sd = {}
    g = torch.Generator().manual_seed(0)
    for key, shape, kind in uo.state_dict_spec(UNET_CFG):
        if kind == "w":
            fan = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan)
        elif kind == "g":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind in ("b", "n"):
            sd[key] = 0.02 * torch.randn(shape, generator=g)
        else:
            sd[key] = torch.tensor(0.5)
    return sd


def run_reference(args, rank):
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    lat = args.size // 8
    inputs = synthetic_host_inputs(args.batch, lat, lat, args.boxes, 0, pin=False)
    ref = ReferenceUNet(cpu_state_dict(), "cpu")
    evals = args.plms_steps + 1
    for _ in range(args.warmup):
        cpu_pair_seconds(ref, inputs, 1.0, n_threads)
    t = []
    for i in range(args.steps):
        t.append(cpu_pair_seconds(ref, inputs, 1.0 if i % 2 == 0 else 0.0, n_threads))
    t_pair = float(np.mean(t))
    value = args.batch / (evals * t_pair)
    sample = (f"each step = 1 [cond, uncond] UNet evaluation pair (2 of the {2 * evals} per image) of the reference's UNetModel "
              f"({ref.kind}) at latent {lat}x{lat}, B={args.batch}, fp32 torch CPU, alternating gate 1/0; "
              f"images/sec = B / ({evals} x mean pair time)")
    line = dict(metric="images_per_sec_512x512_50plms_6boxes" if args.size == 512 else f"images_per_sec_{args.size}",
                impl="reference", value=value, unit="images/s", n_gpus=0, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_pair, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args, 1),
                cpu_baseline=dict(value=value, unit="images/s", cores=n_threads, kind=ref.kind, sample=sample),
                e2e=dict(value=value, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return dict(workload=f"{args.size}x{args.size}, {args.plms_steps} PLMS steps, CFG {GUIDANCE}, {args.boxes} boxes, batch={args.batch} per GPU "
                         f"(BASELINE.json configs[1])" if args.batch == 1 and args.size == 512 else
                         f"{args.size}x{args.size}, {args.plms_steps} PLMS steps, CFG {GUIDANCE}, {args.boxes} boxes, batch={args.batch} per GPU",
                per_gpu_batch=args.batch, global_batch=args.batch * world, latent=args.size // 8, plms_steps=args.plms_steps,
                guidance=GUIDANCE, alpha_type=list(ALPHA_TYPE), boxes=args.boxes, parallelism=f"dp{world}",
                l2="working set (2.5 GB fp16 weights streamed per UNet evaluation) exceeds the 126 MB L2; no flush needed")


# ---------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="images per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: total images per step over all GPUs")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--plms-steps", type=int, default=50)
    ap.add_argument("--boxes", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true", help="sampler only (no VAE decode / image conversion in the step)")
    ap.add_argument("--cpu-baseline-pairs", type=int, default=1)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    scaling = "weak"
    if args.global_batch:
        assert args.global_batch % world == 0
        args.batch, scaling = args.global_batch // world, "strong"
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from functools import partial
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the sm_100a library has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when the first communicator is created; the contract is ONE JSON line
        # on stdout, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    B, lat, S = args.batch, args.size // 8, args.plms_steps
    model = build_model(dev)
    vae = None if args.no_decode else build_vae(dev)
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(dev)
    sampler = PLMSSampler(diffusion, model, alpha_generator_func=partial(alpha_generator, type=list(ALPHA_TYPE)),
                          set_alpha_scale=set_alpha_scale)
    host = synthetic_host_inputs(B, lat, lat, args.boxes, rank, pin=True)
    devin = {k: v.to(dev) for k, v in host.items()}
    from layoutllm_t2i_b200 import shard
    out_host = torch.empty(B, 4, lat, lat).pin_memory() if vae is None else torch.empty(B, 8 * lat, 8 * lat, 3, dtype=torch.uint8).pin_memory()
    decode_events = []

    def one_image_batch(src, from_host):
        t = {k: (v.to(dev, non_blocking=True) if from_host else v) for k, v in src.items()}
        g = model.grounding_tokenizer_input.prepare(dict(boxes=t["boxes"], masks=t["masks"], text_embeddings=t["text_embeddings"]), None)
        inp = dict(x=t["x"].clone(), timesteps=None, context=t["context"], relations=t["relations"], grounding_input=g,
                   inpainting_extra_input=None, grounding_extra_input=None)
        z = sampler.sample(S=S, shape=(B, 4, lat, lat), input=inp, uc=t["uc"], guidance_scale=GUIDANCE)
        if world > 1:
            z_all = shard.gather_latents(z, B * world)    # the path's single collective: final latents over NVLink
        if vae is not None:
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            if from_host:                                 # uint8 HWC images -> pinned host memory, one copy, then sync
                vae.decode_to_uint8(z, out_host, sync=True)
            else:
                vae.engine().decode(z, images_u8=True)
            d1.record()
            decode_events.append((d0, d1))
        elif from_host:
            out_host.copy_(z, non_blocking=True)
            torch.cuda.current_stream().synchronize()     # the caller holds the result on the host
        return z

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            z = one_image_batch(host if from_host else devin, from_host)
        e1.record()
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), z

    for _ in range(args.warmup):
        one_image_batch(devin, False)
    eng = model.engine(30)
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = eng.launch_count + (vae.engine().launch_count if vae is not None else 0)
    decode_events.clear()
    ms_dev, z = timed(False, args.steps)
    launches = eng.launch_count + (vae.engine().launch_count if vae is not None else 0) - l0
    ms_decode = sum(a.elapsed_time(b) for a, b in decode_events)
    ms_e2e, z2 = timed(True, args.steps)
    clk = clocks.stop() if clocks else None
    finite = bool(torch.isfinite(z).all().item())

    # per-kernel-class device times IN THE TIMED MODE: the class brackets are external event-record nodes inside the
    # captured CUDA graphs of the evaluation (ltt_profile_enable(m, 2)); one image batch captures the instrumented
    # graphs, a second one is measured.  Each bracket holds the times of its graph's last replay and is weighted by the
    # graph's replay count, i.e. the per-class sums cover every evaluation of the image batch.
    eng.profile(2)
    one_image_batch(devin, False)
    eng.profile(2)
    torch.cuda.synchronize()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    one_image_batch(devin, False)
    pe1.record()
    torch.cuda.synchronize()
    ms_instrumented = pe0.elapsed_time(pe1)
    prof = eng.profile_report()
    eng.profile(0)

    if rank == 0:
        pk = peaks()
        imgs = B * world * args.steps
        value = imgs / (ms_dev / 1e3)
        e2e_v = imgs / (ms_e2e / 1e3)
        evals = S + 1
        n1 = sum(1 for a in alpha_generator(S) if a != 0) + 1          # +1: the Euler predictor's second evaluation
        gf1, gf0 = GF_FWD.get(lat, GF_FWD[64])
        f_alg_tf = 2 * B * (n1 * gf1 + (evals - n1) * gf0) / 1e3          # cond + uncond, fuser elided at alpha = 0
        if vae is not None:
            f_alg_tf += B * GF_VAE_DECODE_64 * (lat / 64.0) ** 2 / 1e3
        # The event nodes cost the programmatic-dependent-launch overlap between neighbouring kernels, so the
        # instrumented image batch is a little slower than a timed one: every class time is scaled by
        # (timed ms per step) / (instrumented ms per step), which makes the classes sum to <= ms_per_step.
        step_ms = ms_dev / args.steps
        k_scale = step_ms / ms_instrumented
        cls = {k: dict(v, ms_raw=v["ms"], ms=v["ms"] * k_scale) for k, v in prof.items()}
        named = ("gemm_tc", "attn_tc", "groupnorm", "layernorm", "relation")
        other_ms = max(step_ms - sum(cls[k]["ms"] for k in named), 0.0)
        dom = max(named, key=lambda k: cls[k]["ms"])
        d = cls[dom]
        if dom in ("gemm_tc", "attn_tc"):
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roof = dict(bound="tensor", kernel=dom, achieved=ach, peak=pk["tflops"], unit="TFLOP/s", frac=ach / pk["tflops"], traffic=None)
        else:
            ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            roof = dict(bound="hbm", kernel=dom, achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], traffic=None)
        try:      # DRAM bytes per launch of the class, from the committed ncu pass of THIS build over the same workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))
            roof["traffic"] = tr["per_launch_bytes"].get(dom)
            roof["traffic_note"] = tr.get("note", "dram__bytes_read+write per launch, class average over one evaluation pair")
        except Exception:  # noqa: BLE001
            pass
        roof.update(peak_source=pk["source"], launches=d["launches"], avg_launch_us=1e3 * d["ms"] / max(d["launches"], 1),
                    alg_bytes_per_launch=d["bytes"] / max(d["launches"], 1), alg_flops_per_launch=d["flops"] / max(d["launches"], 1),
                    share_of_step=d["ms"] / step_ms,
                    classes=dict({k: dict(ms=round(v["ms"], 3), ms_instrumented=round(v["ms_raw"], 3), launches=v["launches"],
                                          alg_tflop=round(v["flops"] / 1e12, 3), alg_gbytes=round(v["bytes"] / 1e9, 3),
                                          tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] else None))
                                  for k, v in cls.items() if k in named},
                                 other=dict(ms=round(other_ms, 3))),
                    ms_per_step=round(step_ms, 3), ms_per_step_instrumented=round(ms_instrumented, 3),
                    note="class times: external CUDA-event nodes inside the replayed graphs of one more image batch of the same "
                         "workload (last replay x replay count per graph), scaled by ms_per_step / ms_per_step_instrumented; "
                         "classes + other = ms_per_step")
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = dict(metric="images_per_sec_512x512_50plms_6boxes" if (args.size == 512 and S == 50 and args.boxes == 6) else
                           f"images_per_sec_{args.size}x{args.size}_{S}plms_{args.boxes}boxes",
                    value=value, unit="images/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="f16",
                    data="synthetic (seeded noise/text/box tensors, random-init weights of the LayoutLLM-T2I UNet)",
                    config=workload_config(args, world),
                    e2e=dict(value=e2e_v, unit="images/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=out_host.numel() * out_host.element_size(), ms_per_step=ms_e2e / args.steps,
                             result="uint8 HWC images" if vae is not None else "final latents (fp32)"),
                    includes_vae_decode=vae is not None,
                    sampler_only=dict(value=imgs / ((ms_dev - ms_decode) / 1e3), unit="images/s",
                                      ms_per_step=(ms_dev - ms_decode) / args.steps, decode_ms_per_step=ms_decode / args.steps),
                    gpu_launches=int(launches), clocks=clk, roofline=roof,
                    step_tensor_roofline=dict(alg_tflop_per_step=f_alg_tf, achieved_tflops=f_alg_tf * args.steps / (ms_dev / 1e3),
                                              frac=f_alg_tf * args.steps / (ms_dev / 1e3) / pk["tflops"], peak=pk["tflops"]),
                    output_finite=finite)
        if world == 1 and not args.no_cpu_baseline:
            n_threads = os.cpu_count() or 1
            sd = cpu_state_dict()
            hin = {k: v.clone() for k, v in synthetic_host_inputs(B, lat, lat, args.boxes, 0, pin=False).items()}
            ref = ReferenceUNet(sd, "cpu")
            tp = [cpu_pair_seconds(ref, hin, 1.0, n_threads) for _ in range(args.cpu_baseline_pairs)]
            line["cpu_baseline"] = dict(value=B / (evals * float(np.mean(tp))), unit="images/s", cores=n_threads, kind=ref.kind,
                                        sample=f"{len(tp)} [cond, uncond] UNet evaluation pair(s) of the reference's UNetModel ({ref.kind}; fp32, "
                                               f"torch CPU, {n_threads} threads) at latent {lat}x{lat}, B={B}; extrapolated x{evals} pairs per image")
            try:      # the same modules as eager fp16-autocast PyTorch on this GPU (context for the ">= 6x eager GPU" target)
                if ref.kind == "reference":
                    ref.model.to(dev)
                    ref.dev = dev
                else:
                    ref = ReferenceUNet(sd, dev)
                ms1 = gpu_eager_pair_ms(ref, hin, 1.0)
                ms0 = gpu_eager_pair_ms(ref, hin, 0.0)
                line["reference_gpu_eager"] = dict(
                    value=B / ((n1 * ms1 + (evals - n1) * ms0) / 1e3), unit="images/s", kind=ref.kind, dtype="fp16 autocast",
                    ms_per_pair=dict(gate1=ms1, gate0=ms0),
                    sample=f"3 [cond, uncond] evaluation pairs per gate value of the reference's UNetModel ({ref.kind}) as eager PyTorch on the "
                           f"same GPU (two separate B={B} forwards per pair, as the reference's sampler does); extrapolated to {evals} pairs per image")
            except Exception as ex:  # noqa: BLE001
                line["reference_gpu_eager"] = dict(unavailable=repr(ex)[:200])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
