"""Representative single launches of the tensor-core kernels for `ncu --set full --profile-from-start off`:
every op is warmed up, then launched ONCE between cudaProfilerStart/Stop.
usage: ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/x python tools/ncu_ops.py [names...]
names: geglu0 conv0 smallk lin0 attn40 attn80 attn160 xattn"""
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
names = sys.argv[1:] or ["geglu0", "conv0", "smallk", "lin0", "attn40", "attn80"]


def linear_fn(M, N, K, act=0, res=False):
    a = oc.rn(M, K, dtype=torch.float16)
    w = oc.rn(N, K, seed=1, scale=1 / math.sqrt(K)).half()
    b = oc.rn(N, seed=2, scale=0.1)
    nout = N // 2 if act == 2 else N
    r = oc.rn(M, nout, seed=3, dtype=torch.float16) if res else None
    return lambda: oc.linear(a, w, b, act, r)


def conv_fn(B, H, W, C, N):
    x = oc.rn(B, H, W, C, dtype=torch.float16)
    wp = oc.rn(N, 9 * C, seed=1, scale=1 / math.sqrt(9 * C)).half()
    b = oc.rn(N, seed=2, scale=0.1)
    out = torch.empty(B, H, W, N, device=DEV, dtype=torch.float16)
    return lambda: L.check(L.lib().ltt_op_conv3x3(L.ptr(x), B, H, W, C, L.ptr(wp), N, L.ptr(b), None, L.ptr(out), L.stream_ptr()), "conv")


def attn_fn(B, heads, d, nq, nk):
    q, k, v, qp, kp, vt, rows_k, pitch = oc.attention_inputs(B, heads, d, nq, nk)
    C = heads * d
    out = torch.zeros(B, nq, C, device=DEV, dtype=torch.float16)
    return lambda: L.check(L.lib().ltt_op_attention(L.ptr(qp), nq, L.ptr(kp), rows_k, L.ptr(vt), pitch, B, heads, d, oc.dpad_of(d), nq,
                                                    nk, d ** -0.5, L.ptr(out), C, L.stream_ptr()), "attention")


def gn_fn(B, HW, c0, c1):
    x0 = oc.rn(B, HW, c0, dtype=torch.float16)
    x1 = oc.rn(B, HW, c1, seed=1, dtype=torch.float16) if c1 else None
    C = c0 + c1
    g = 1 + 0.1 * oc.rn(C, seed=2)
    b = 0.1 * oc.rn(C, seed=3)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    return lambda: L.check(L.lib().ltt_op_groupnorm(L.ptr(x0), c0, L.ptr(x1), c1, B, HW, L.ptr(g), L.ptr(b), 1e-5, 1, L.ptr(out), L.stream_ptr()), "gn")


def small_attn_fn(B, d):
    C = 8 * d
    q = oc.rn(B, 30, C, dtype=torch.float16)
    k = oc.rn(B, 10, C, seed=1, dtype=torch.float16)
    v = oc.rn(B, 10, C, seed=2, dtype=torch.float16)
    out = torch.empty(B, 30, C, device=DEV, dtype=torch.float16)
    return lambda: L.check(L.lib().ltt_op_small_attention(L.ptr(q), L.ptr(k), L.ptr(v), B, 30, 10, 8, d, d ** -0.5, L.ptr(out), L.stream_ptr()), "sa")


OPS = {
    "gn0": lambda: gn_fn(2, 4096, 320, 0),
    "gn960": lambda: gn_fn(2, 4096, 640, 320),
    "sattn": lambda: small_attn_fn(1, 160),
    "geglu0": lambda: linear_fn(8192, 2560, 320, act=2),
    "geglu1": lambda: linear_fn(2048, 5120, 640, act=2),
    "conv0": lambda: conv_fn(2, 64, 64, 320, 320),
    "conv1": lambda: conv_fn(2, 32, 32, 640, 640),
    "smallk": lambda: linear_fn(512, 1280, 1280, res=True),
    "lin0": lambda: linear_fn(8192, 320, 320, res=True),
    "ff2_0": lambda: linear_fn(8192, 320, 1280, res=True),
    "attn40": lambda: attn_fn(2, 8, 40, 4096, 4126),
    "attn80": lambda: attn_fn(2, 8, 80, 1024, 1054),
    "attn160": lambda: attn_fn(2, 8, 160, 256, 286),
    "xattn": lambda: attn_fn(2, 8, 40, 4096, 77),
}
fns = [(n, OPS[n]()) for n in names]
for _, f in fns:
    for _ in range(3):
        f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for n, f in fns:
    f()
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled:", " ".join(names))
