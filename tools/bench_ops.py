"""Micro-benchmark of the tensor-core kernels on the UNet's layer shapes (B=2 = one [cond ; uncond] pair at 64x64).
Each op is launched `iters` times back to back and timed with CUDA events; inputs rotate over enough buffers to defeat L2
for the weight-streaming shapes."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import ops_checks as oc  # noqa: E402
from layoutllm_t2i_b200 import _lib as L  # noqa: E402

DEV = "cuda"
only = sys.argv[1] if len(sys.argv) > 1 else ""
iters = 20


def timeit(fn):
    """us per launch: `iters` launches captured into one CUDA graph (no Python / launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * iters) * 1e3   # us


def bench_linear(M, N, K, act=0, res=False):
    a = oc.rn(M, K, dtype=torch.float16)
    w = oc.rn(N, K, seed=1, scale=1 / math.sqrt(K)).half()
    b = oc.rn(N, seed=2, scale=0.1)
    nout = N // 2 if act == 2 else N
    r = oc.rn(M, nout, seed=3, dtype=torch.float16) if res else None
    us = timeit(lambda: oc.linear(a, w, b, act, r))
    fl = 2.0 * M * N * K
    print(f"linear {M:5d}x{N:5d}x{K:5d} act={act} res={int(res)}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
          f"(weights {N * K * 2 / us / 1e3:6.1f} GB/s)", flush=True)


def bench_conv(B, H, W, C, N):
    x = oc.rn(B, H, W, C, dtype=torch.float16)
    wp = oc.rn(N, 9 * C, seed=1, scale=1 / math.sqrt(9 * C)).half()
    b = oc.rn(N, seed=2, scale=0.1)
    out = torch.empty(B, H, W, N, device=DEV, dtype=torch.float16)
    fn = lambda: L.check(L.lib().ltt_op_conv3x3(L.ptr(x), B, H, W, C, L.ptr(wp), N, L.ptr(b), None, L.ptr(out), L.stream_ptr()), "conv")
    us = timeit(fn)
    fl = 2.0 * B * H * W * N * 9 * C
    print(f"conv3x3 {B}x{H}x{W} {C:4d}->{N:4d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  (weights {N * 9 * C * 2 / us / 1e3:6.1f} GB/s)", flush=True)


def bench_attn(B, heads, d, nq, nk):
    q, k, v, qp, kp, vt, rows_k, pitch = oc.attention_inputs(B, heads, d, nq, nk)
    C = heads * d
    out = torch.zeros(B, nq, C, device=DEV, dtype=torch.float16)
    fn = lambda: L.check(L.lib().ltt_op_attention(L.ptr(qp), nq, L.ptr(kp), rows_k, L.ptr(vt), pitch, B, heads, d, oc.dpad_of(d), nq,
                                                  nk, d ** -0.5, L.ptr(out), C, L.stream_ptr()), "attention")
    us = timeit(fn)
    fl = 4.0 * B * heads * nq * nk * d
    print(f"attention B={B} d={d:3d} nq={nq:5d} nk={nk:5d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)


def bench_gn(B, HW, c0, c1):
    x0 = oc.rn(B, HW, c0, dtype=torch.float16)
    x1 = oc.rn(B, HW, c1, seed=1, dtype=torch.float16) if c1 else None
    C = c0 + c1
    g = 1 + 0.1 * oc.rn(C, seed=2)
    b = 0.1 * oc.rn(C, seed=3)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    fn = lambda: L.check(L.lib().ltt_op_groupnorm(L.ptr(x0), c0, L.ptr(x1), c1, B, HW, L.ptr(g), L.ptr(b), 1e-5, 1, L.ptr(out), L.stream_ptr()), "gn")
    us = timeit(fn)
    print(f"groupnorm {B}x{HW} {c0}+{c1}: {us:8.1f} us  {B * HW * C * 4 / us / 1e3:7.1f} GB/s (read+write)", flush=True)


def bench_ln(M, C, dtype=torch.float16):
    x = oc.rn(M, C).to(dtype)
    g = 1 + 0.1 * oc.rn(C, seed=1)
    b = 0.1 * oc.rn(C, seed=2)
    o16 = torch.empty(M, C, device=DEV, dtype=torch.float16)
    fn = lambda: L.check(L.lib().ltt_op_layernorm(L.ptr(x), 0 if dtype == torch.float16 else 1, M, C, L.ptr(g), L.ptr(b), 1e-5, L.ptr(o16), None, L.stream_ptr()), "ln")
    us = timeit(fn)
    print(f"layernorm {M}x{C} {str(dtype)[6:]}: {us:8.1f} us  {M * C * (x.element_size() + 2) / us / 1e3:7.1f} GB/s", flush=True)


def bench_small_attn(B, d):
    C = 8 * d
    q = oc.rn(B, 30, C, dtype=torch.float16)
    k = oc.rn(B, 10, C, seed=1, dtype=torch.float16)
    v = oc.rn(B, 10, C, seed=2, dtype=torch.float16)
    out = torch.empty(B, 30, C, device=DEV, dtype=torch.float16)
    fn = lambda: L.check(L.lib().ltt_op_small_attention(L.ptr(q), L.ptr(k), L.ptr(v), B, 30, 10, 8, d, d ** -0.5, L.ptr(out), L.stream_ptr()), "sa")
    print(f"relation attention B={B} d={d}: {timeit(fn):8.1f} us", flush=True)


print(torch.cuda.get_device_name(0))
if only == "small":
    for B, HW, c0, c1 in ((2, 4096, 320, 0), (2, 4096, 640, 320), (2, 4096, 320, 320), (2, 1024, 640, 0), (2, 1024, 1280, 640),
                          (2, 256, 1280, 0), (2, 256, 1280, 1280), (2, 64, 1280, 0), (2, 64, 1280, 1280)):
        bench_gn(B, HW, c0, c1)
    for M, C in ((8192, 320), (2048, 640), (512, 1280), (128, 1280), (30, 1280)):
        bench_ln(M, C)
    bench_ln(8192, 320, torch.float32)
    for d in (40, 80, 160):
        bench_small_attn(1, d)
if only == "norms":      # memory-bound kernels across batch sizes (B = 2 is one image's [cond ; uncond] pair)
    for B in (2, 16, 128):
        for HW, c0, c1 in ((4096, 320, 0), (4096, 640, 320), (1024, 640, 0), (256, 1280, 0)):
            bench_gn(B, HW, c0, c1)
        for N, C in ((4096, 320), (1024, 640), (256, 1280)):
            bench_ln(B * N, C)
        bench_ln(B * 4096, 320, torch.float32)
if not only or only == "linear":
    for M, C in ((8192, 320), (2048, 640), (512, 1280), (128, 1280)):
        bench_linear(M, C, C)                 # proj / to_out
        bench_linear(M, C, C, res=True)
        bench_linear(M, 3 * C, C)             # fused qkv width
        bench_linear(M, 8 * C, C, act=2)      # GEGLU ff1
        bench_linear(M, C, 4 * C, res=True)   # ff2
    bench_linear(60, 1280, 1280)
    bench_linear(60, 10240, 1280, act=2)
    bench_linear(154, 2560, 768)
if only == "peak":      # large square-ish GEMMs: the kernel's own main-loop ceiling per tile width / pair mode
    bench_linear(16384, 4096, 4096)
    bench_linear(16384, 5120, 2560)
    bench_conv(16, 64, 64, 320, 320)
    bench_conv(16, 32, 32, 640, 640)
if not only or only == "conv":
    bench_conv(2, 64, 64, 320, 320)
    bench_conv(2, 64, 64, 640, 320)
    bench_conv(2, 32, 32, 640, 640)
    bench_conv(2, 32, 32, 1280, 640)
    bench_conv(2, 16, 16, 1280, 1280)
    bench_conv(2, 16, 16, 2560, 1280)
    bench_conv(2, 8, 8, 1280, 1280)
    bench_conv(16, 64, 64, 320, 320)
if not only or only == "attn":
    bench_attn(2, 8, 40, 4096, 4096)
    bench_attn(2, 8, 40, 4096, 4126)
    bench_attn(2, 8, 80, 1024, 1054)
    bench_attn(2, 8, 160, 256, 286)
    bench_attn(2, 8, 40, 4096, 77)
    bench_attn(16, 8, 40, 4096, 4126)
