// VAE decode + image post-processing on the sm_100a kernels.
//
// Mirrors (reference, /root/reference/GLIGEN/ldm): AutoencoderKL.decode models/autoencoder.py:40-44
// (z / scale_factor -> post_quant_conv -> Decoder), Decoder.forward modules/diffusionmodules/model.py:538-568,
// ResnetBlock :82-143, AttnBlock :150-202 (single head, d = C, all tokens), Upsample :43-59, Normalize :38-39
// (GroupNorm 32 groups, eps 1e-6); and the callers' post-processing txt2img.py:320-323 / GLIGEN/interface.py:541-545
// (clamp to [-1, 1], * 0.5 + 0.5, * 255, uint8, HWC).
//
// Same building blocks as the UNet: every 3x3 / 1x1 convolution is the tcgen05 implicit GEMM (gemm_tc.cu) on NHWC
// fp16 activations, GroupNorm + swish is the cluster kernel of small_ops.cu (swish applied to the fp32 normalised value,
// where CUDA autocast leaves it), the ResnetBlock's 1x1 nin_shortcut rides along as extra K columns of conv2.  The one
// 4096-token, d = 512 single-head attention block runs as two GEMMs around a row softmax (scores materialised in fp16 as
// torch.bmm does under autocast; 34 GFLOP per image, once per image).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ltt_b200.h"
#include "ltt_ops.h"
#include "ltt_ptx.cuh"

namespace ltt {

#define RCV(expr)               \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)

struct VParam { float* dev = nullptr; std::vector<int64_t> shape; size_t numel = 0; };
struct VNorm { const float* g = nullptr; const float* b = nullptr; };
struct VLin { __half* w = nullptr; const float* bias = nullptr; int N = 0, K = 0; };
struct VRes { std::string p; int cin = 0, cout = 0; bool has_skip = false; VNorm n1, n2; VLin conv1, conv2; };
struct VUp { int C = 0; VLin conv; };

// ---- kernels private to the decoder
__device__ __forceinline__ float r16v(float x) { return __half2float(__float2half_rn(x)); }

// z' = post_quant_conv(z / scale_factor): 1x1 conv E -> Z channels on NCHW fp32 (autocast: fp16 inputs / weights, fp32
// accumulate, fp16 result); written as fp32 NCHW holding fp16 values for the first-conv kernel.
__global__ void vae_prep_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias,
                                float inv_scale, int B, int E, int Z, int HW, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / HW, p = i - b * HW;
        float v[8];
        for (int e = 0; e < E; ++e) v[e] = r16v(z[(b * E + e) * HW + p] * inv_scale);
        for (int c = 0; c < Z; ++c) {
            float acc = 0.f;
            for (int e = 0; e < E; ++e) acc += v[e] * r16v(w[c * E + e]);
            out[(b * Z + c) * HW + p] = r16v(acc + bias[c]);
        }
    }
}

// P[r, :] = fp16(softmax(fp16(S[r, :] * scale))) in place: the reference rounds the scaled logits to fp16 (fp16 tensor
// times a Python scalar), softmax runs in fp32 under autocast and its output is cast to fp16 by the following bmm
// (model.py:186-195).  One CTA per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ s, int n, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[8];
    __half* row = s + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int i = tid * 8; i < n; i += 256 * 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + i);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            mx = fmaxf(mx, fmaxf(r16v(f.x * scale), r16v(f.y * scale)));
        }
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid * 8; i < n; i += 256 * 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + i);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            sum += __expf(r16v(f.x * scale) - mx) + __expf(r16v(f.y * scale) - mx);
        }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w];
    const float inv = 1.0f / sum;
    for (int i = tid * 8; i < n; i += 256 * 8) {
        uint4 u = *reinterpret_cast<const uint4*>(row + i);
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            h[k] = __floats2half2_rn(__expf(r16v(f.x * scale) - mx) * inv, __expf(r16v(f.y * scale) - mx) * inv);
        }
        *reinterpret_cast<uint4*>(row + i) = u;
    }
}

// conv_out: Conv2d(C -> Cout <= 4, 3x3, pad 1) on NHWC fp16 (already GroupNorm + swish) with the callers' image
// post-processing fused: img_f32 (NCHW fp32, fp16-rounded values, may be null) = the decoder output the reference
// returns; img_u8 (NHWC uint8, may be null) = uint8((clamp(v, -1, 1) * 0.5 + 0.5) * 255) as txt2img.py:320-323.
// One warp per output pixel (lanes split the channels in 16-byte vectors), weights [Cout][9][C] fp16 in shared memory.
constexpr int VO_WARPS = 8, VO_PIX = 4;
__global__ void __launch_bounds__(VO_WARPS * 32) vae_conv_out_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                                    const float* __restrict__ bias, int B, int H, int W, int C,
                                                                    int Cout, float* __restrict__ img_f32,
                                                                    uint8_t* __restrict__ img_u8) {
    pdl_launch_dependents();
    extern __shared__ uint4 smem_vo[];
    const int nw = Cout * 9 * C / 8;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) smem_vo[i] = reinterpret_cast<const uint4*>(w)[i];
    pdl_wait();
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = C >> 3;
    const size_t total = (size_t)B * H * W;
    for (int pp = 0; pp < VO_PIX; ++pp) {
        const size_t pix = ((size_t)blockIdx.x * VO_WARPS + warp) * VO_PIX + pp;
        if (pix >= total) break;
        const int b = (int)(pix / ((size_t)H * W));
        const int rem = (int)(pix - (size_t)b * H * W), y = rem / W, xx = rem - y * W;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cv = lane; cv < V; cv += 32) {
            uint4 xv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = y + t / 3 - 1, xs = xx + t % 3 - 1;
                const bool in = yy >= 0 && yy < H && xs >= 0 && xs < W;
                xv[t] = in ? *reinterpret_cast<const uint4*>(x + (((size_t)b * H + yy) * W + xs) * C + cv * 8) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const __half2* xh = reinterpret_cast<const __half2*>(&xv[t]);
#pragma unroll
                for (int co = 0; co < 4; ++co)
                    if (co < Cout) {
                        const uint4 wv = smem_vo[(co * 9 + t) * V + cv];
                        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 a = __half22float2(xh[k]), ww = __half22float2(wh[k]);
                            acc[co] += a.x * ww.x + a.y * ww.y;
                        }
                    }
            }
        }
#pragma unroll
        for (int co = 0; co < 4; ++co) {
#pragma unroll
            for (int o = 16; o; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
        }
        if (lane < Cout && lane < 4) {
            const float a = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
            const float v = r16v(a + bias[lane]);
            if (img_f32) img_f32[(((size_t)b * Cout + lane) * H + y) * W + xx] = v;
            if (img_u8) {
                const float u = (fminf(fmaxf(v, -1.0f), 1.0f) * 0.5f + 0.5f) * 255.0f;
                img_u8[pix * Cout + lane] = (uint8_t)u;      // truncation, as numpy's astype(uint8)
            }
        }
    }
}

}  // namespace ltt

using namespace ltt;

struct ltt_vae {
    ltt_vae_config cfg;
    int device = 0, sms = 148;
    std::map<std::string, VParam> params;
    bool finalized = false;
    std::vector<void*> wptrs, cptrs;
    std::vector<std::string> packed;    // fp32 weights that live on as packed fp16 copies: released after ltt_vae_finalize
    // plan
    std::vector<VRes> res;              // mid.block_1, mid.block_2, then the up blocks in execution order
    std::vector<std::vector<int>> level_res;   // per executed level (high -> low index): indices into res
    std::vector<VUp> ups;               // one per executed level that upsamples
    VNorm attn_norm, norm_out;
    VLin attn_qkv, attn_out;
    int Cmid = 0;
    const float *conv_in_w = nullptr, *conv_in_b = nullptr, *pq_w = nullptr, *pq_b = nullptr;
    __half* out_w = nullptr;
    const float* out_b = nullptr;
    // workspace for (B, h, w)
    int B = 0, h = 0, w = 0;
    float* zq = nullptr;
    __half *act0 = nullptr, *act1 = nullptr, *tnorm = nullptr, *hbuf = nullptr;
    __half *qb = nullptr, *kb = nullptr, *vtb = nullptr, *sbuf = nullptr, *ao = nullptr;
    double* gn_stats = nullptr;
    int64_t launches = 0;
};

namespace ltt {

static int valloc(std::vector<void*>& arena, void** out, size_t bytes) {
    void* p = nullptr;
    LTT_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 16));
    arena.push_back(p);
    *out = p;
    return 0;
}
static void vrelease(std::vector<void*>& arena) {
    for (void* p : arena) cudaFree(p);
    arena.clear();
}
static const VParam* vfind(ltt_vae* v, const std::string& key) {
    auto it = v->params.find(key);
    if (it == v->params.end()) {
        set_error("vae: missing parameter '%s' (load_state_dict incomplete)", key.c_str());
        return nullptr;
    }
    return &it->second;
}
#define VGET(var, key)                    \
    const VParam* var = vfind(v, (key));  \
    if (!var) return -6;

static int vnorm(ltt_vae* v, const std::string& p, VNorm* n) {
    VGET(g, p + ".weight")
    VGET(b, p + ".bias")
    n->g = g->dev; n->b = b->dev;
    return 0;
}
// 3x3 conv [cout, cin, 3, 3] (+ optional 1x1 skip [cout, cskip, 1, 1] appended along K) -> packed fp16 [cout, K]
static int vconv3(ltt_vae* v, const std::string& p, int cin, int cout, const std::string& skip, int cskip, VLin* out) {
    VGET(w, p + ".weight")
    VGET(b, p + ".bias")
    const int K = 9 * cin + cskip;
    void* q;
    RCV(valloc(v->wptrs, &q, (size_t)cout * K * 2));
    RCV(pack_conv_launch(w->dev, cout, cin, 9, 0, cin, (__half*)q, K, 0, 0));
    v->packed.push_back(p + ".weight");
    const float* bias = b->dev;
    if (cskip) {
        VGET(sw, skip + ".weight")
        VGET(sb, skip + ".bias")
        RCV(pack_conv_launch(sw->dev, cout, cskip, 1, 0, cskip, (__half*)q, K, 9 * cin, 0));
        v->packed.push_back(skip + ".weight");
        std::vector<float> hb(cout), hs(cout);
        LTT_CUDA_OK(cudaMemcpy(hb.data(), b->dev, cout * 4, cudaMemcpyDeviceToHost));
        LTT_CUDA_OK(cudaMemcpy(hs.data(), sb->dev, cout * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < cout; ++i) hb[i] += hs[i];
        void* cb;
        RCV(valloc(v->wptrs, &cb, cout * 4));
        LTT_CUDA_OK(cudaMemcpy(cb, hb.data(), cout * 4, cudaMemcpyHostToDevice));
        bias = (const float*)cb;
    }
    *out = VLin{(__half*)q, bias, cout, K};
    return 0;
}
static int vres(ltt_vae* v, const std::string& p, int cin, int cout) {
    VRes r;
    r.p = p; r.cin = cin; r.cout = cout; r.has_skip = cin != cout;
    RCV(vnorm(v, p + ".norm1", &r.n1));
    RCV(vnorm(v, p + ".norm2", &r.n2));
    RCV(vconv3(v, p + ".conv1", cin, cout, "", 0, &r.conv1));
    RCV(vconv3(v, p + ".conv2", cout, cout, p + ".nin_shortcut", r.has_skip ? cin : 0, &r.conv2));
    v->res.push_back(r);
    return 0;
}

// Decoder.__init__ (model.py:462-536)
static int vae_build(ltt_vae* v) {
    const ltt_vae_config& c = v->cfg;
    vrelease(v->wptrs);
    v->packed.clear();
    v->res.clear(); v->level_res.clear(); v->ups.clear();
    int block_in = c.ch * c.ch_mult[c.n_levels - 1];
    v->Cmid = block_in;
    {
        VGET(w, "decoder.conv_in.weight")
        VGET(b, "decoder.conv_in.bias")
        VGET(pw, "post_quant_conv.weight")
        VGET(pb, "post_quant_conv.bias")
        v->conv_in_w = w->dev; v->conv_in_b = b->dev; v->pq_w = pw->dev; v->pq_b = pb->dev;
    }
    RCV(vres(v, "decoder.mid.block_1", block_in, block_in));
    RCV(vres(v, "decoder.mid.block_2", block_in, block_in));
    RCV(vnorm(v, "decoder.mid.attn_1.norm", &v->attn_norm));
    {   // q / k / v 1x1 convs stacked into one [3C, C] matrix with their biases
        const int C = block_in;
        void *q, *bb;
        RCV(valloc(v->wptrs, &q, (size_t)3 * C * C * 2));
        RCV(valloc(v->wptrs, &bb, (size_t)3 * C * 4));
        const char* names[3] = {"q", "k", "v"};
        for (int i = 0; i < 3; ++i) {
            VGET(w, std::string("decoder.mid.attn_1.") + names[i] + ".weight")
            VGET(b, std::string("decoder.mid.attn_1.") + names[i] + ".bias")
            RCV(pack_rows_launch(w->dev, C, C, (__half*)q, i * C, 0, 0));
            v->packed.push_back(std::string("decoder.mid.attn_1.") + names[i] + ".weight");
            LTT_CUDA_OK(cudaMemcpy((float*)bb + i * C, b->dev, C * 4, cudaMemcpyDeviceToDevice));
        }
        v->attn_qkv = VLin{(__half*)q, (const float*)bb, 3 * C, C};
        VGET(w, "decoder.mid.attn_1.proj_out.weight")
        VGET(b, "decoder.mid.attn_1.proj_out.bias")
        void* o;
        RCV(valloc(v->wptrs, &o, (size_t)C * C * 2));
        RCV(pack_rows_launch(w->dev, C, C, (__half*)o, 0, 0, 0));
        v->packed.push_back("decoder.mid.attn_1.proj_out.weight");
        v->attn_out = VLin{(__half*)o, b->dev, C, C};
    }
    char buf[128];
    for (int lvl = c.n_levels - 1; lvl >= 0; --lvl) {
        const int block_out = c.ch * c.ch_mult[lvl];
        std::vector<int> idx;
        for (int i = 0; i <= c.num_res_blocks; ++i) {
            snprintf(buf, sizeof(buf), "decoder.up.%d.block.%d", lvl, i);
            RCV(vres(v, buf, block_in, block_out));
            idx.push_back((int)v->res.size() - 1);
            block_in = block_out;
        }
        v->level_res.push_back(idx);
        if (lvl != 0) {
            snprintf(buf, sizeof(buf), "decoder.up.%d.upsample.conv", lvl);
            VUp u;
            u.C = block_in;
            RCV(vconv3(v, buf, block_in, block_in, "", 0, &u.conv));
            v->ups.push_back(u);
        }
    }
    RCV(vnorm(v, "decoder.norm_out", &v->norm_out));
    {
        VGET(w, "decoder.conv_out.weight")
        VGET(b, "decoder.conv_out.bias")
        void* q;
        RCV(valloc(v->wptrs, &q, (size_t)c.out_ch * 9 * block_in * 2));
        RCV(pack_conv_launch(w->dev, c.out_ch, block_in, 9, 0, block_in, (__half*)q, 9 * block_in, 0, 0));
        v->packed.push_back("decoder.conv_out.weight");
        v->out_w = (__half*)q; v->out_b = b->dev;
    }
    LTT_CUDA_OK(cudaDeviceSynchronize());
    for (const std::string& k : v->packed) {       // (a later finalize needs the state_dict loaded again, as load_state_dict does)
        auto it = v->params.find(k);
        if (it != v->params.end()) {
            cudaFree(it->second.dev);
            v->params.erase(it);
        }
    }
    v->packed.clear();
    return 0;
}

static int vae_workspace(ltt_vae* v, int B, int h, int w) {
    vrelease(v->cptrs);
    const ltt_vae_config& c = v->cfg;
    size_t max_act = 0;
    {
        int hh = h, ww = w, ch = v->Cmid;
        max_act = (size_t)B * hh * ww * ch;
        size_t li = 0;
        for (int lvl = c.n_levels - 1; lvl >= 0; --lvl, ++li) {
            for (int ri : v->level_res[li]) {
                max_act = std::max(max_act, (size_t)B * hh * ww * std::max(v->res[ri].cin, v->res[ri].cout));
                ch = v->res[ri].cout;
            }
            if (lvl != 0) {
                hh *= 2; ww *= 2;
                max_act = std::max(max_act, (size_t)B * hh * ww * ch);
            }
        }
    }
    const size_t N = (size_t)h * w, C = v->Cmid;
    auto A = [&](auto** p, size_t bytes) { return valloc(v->cptrs, (void**)p, bytes); };
    RCV(A(&v->zq, (size_t)B * c.z_channels * N * 4));
    RCV(A(&v->act0, max_act * 2)); RCV(A(&v->act1, max_act * 2)); RCV(A(&v->tnorm, max_act * 2)); RCV(A(&v->hbuf, max_act * 2));
    RCV(A(&v->qb, (size_t)B * N * C * 2)); RCV(A(&v->kb, (size_t)B * N * C * 2)); RCV(A(&v->vtb, (size_t)B * N * C * 2));
    RCV(A(&v->ao, (size_t)B * N * C * 2));
    RCV(A(&v->sbuf, N * N * 2));
    RCV(A(&v->gn_stats, (size_t)B * 64 * sizeof(double)));
    v->B = B; v->h = h; v->w = w;
    return 0;
}

struct VRun {
    ltt_vae* v;
    cudaStream_t st;
    int gemm(int B, int H, int W, int N, std::initializer_list<GemmSrc> srcs, const __half* w, int Kw, const GemmEpilogue& epi,
             int w_static) {
        GemmProblem p{};
        p.B = B; p.H = H; p.W = W; p.N = N; p.nsrc = 0; p.Ktot = 0;
        for (auto& s : srcs) {
            p.src[p.nsrc++] = s;
            p.Ktot += s.taps * s.channels;
        }
        if (p.Ktot != Kw) {
            set_error("vae: internal GEMM shape mismatch K=%d/%d", p.Ktot, Kw);
            return -7;
        }
        p.w = w; p.w_static = w_static; p.epi = epi;
        v->launches++;
        return gemm_tc_launch(p, v->sms, st);
    }
    int gn(const __half* x, int C, int B, int HW, const VNorm& n, int silu, __half* out) {
        const int rc = groupnorm_fused_launch(x, C, C, nullptr, 0, 0, B, HW, 32, n.g, n.b, 1e-6f, silu, out, st);
        if (rc <= 0) {
            v->launches += 1;
            return rc;
        }
        RCV(gn_stats_launch(x, C, C, nullptr, 0, 0, B, HW, 32, v->gn_stats, st));
        RCV(gn_apply_launch(x, C, C, nullptr, 0, 0, B, HW, 32, v->gn_stats, n.g, n.b, 1e-6f, silu, out, st));
        v->launches += 3;
        return 0;
    }
};

static GemmEpilogue vepi(void* out, int ldo, const float* bias) {
    GemmEpilogue e;
    e.out = out; e.ldo = ldo; e.out_dtype = DT_F16; e.bias = bias;
    return e;
}

// ResnetBlock.forward with temb = None (model.py:123-143)
static int vae_res(VRun& r, const VRes& w, const __half* x, int B, int H, int W, __half* out) {
    ltt_vae* v = r.v;
    RCV(r.gn(x, w.cin, B, H * W, w.n1, 2, v->tnorm));
    RCV(r.gemm(B, H, W, w.cout, {GemmSrc{v->tnorm, w.cin, w.cin, 9}}, w.conv1.w, w.conv1.K, vepi(v->hbuf, w.cout, w.conv1.bias), 1));
    RCV(r.gn(v->hbuf, w.cout, B, H * W, w.n2, 2, v->tnorm));
    GemmEpilogue e = vepi(out, w.cout, w.conv2.bias);
    if (w.has_skip) {
        RCV(r.gemm(B, H, W, w.cout, {GemmSrc{v->tnorm, w.cout, w.cout, 9}, GemmSrc{x, w.cin, w.cin, 1}}, w.conv2.w, w.conv2.K, e, 1));
    } else {
        e.res = x; e.res_dtype = DT_F16; e.ldr = w.cin;
        RCV(r.gemm(B, H, W, w.cout, {GemmSrc{v->tnorm, w.cout, w.cout, 9}}, w.conv2.w, w.conv2.K, e, 1));
    }
    return 0;
}

// AttnBlock.forward (model.py:176-202)
static int vae_attn(VRun& r, const __half* x, int B, int H, int W, __half* out) {
    ltt_vae* v = r.v;
    const int N = H * W, C = v->Cmid;
    if (N % 64) {
        set_error("vae: attention block needs h*w %% 64 == 0 (h*w = %d)", N);
        return -1;
    }
    RCV(r.gn(x, C, B, N, v->attn_norm, 0, v->tnorm));
    {
        GemmEpilogue e;
        e.bias = v->attn_qkv.bias;
        e.out_mode = OUT_QKV;
        e.q = v->qb; e.k = v->kb; e.vt = v->vtb;
        e.C = C; e.dhead = C; e.dpad = C; e.rows_q = N; e.rows_k = N; e.pitch_v = N; e.tokens = N; e.qkv_base = 0;
        RCV(r.gemm(B, H, W, 3 * C, {GemmSrc{v->tnorm, C, C, 1}}, v->attn_qkv.w, C, e, 1));
    }
    const float scale = 1.0f / sqrtf((float)C);
    for (int b = 0; b < B; ++b) {
        // S = q k^T (fp16, as torch.bmm under autocast); the "weights" of this GEMM are this sample's keys
        RCV(r.gemm(1, 1, N, N, {GemmSrc{v->qb + (size_t)b * N * C, C, C, 1}}, v->kb + (size_t)b * N * C, C, vepi(v->sbuf, N, nullptr), 0));
        LTT_CUDA_OK(launch_k(softmax_rows_kernel, dim3(N), dim3(256), 0, r.st, v->sbuf, N, scale));
        v->launches++;
        // h_[i, c] = sum_j P[i, j] v[j, c]: B operand = V^T [C, N], written that way by the QKV epilogue
        RCV(r.gemm(1, 1, N, C, {GemmSrc{v->sbuf, N, N, 1}}, v->vtb + (size_t)b * C * N, N, vepi(v->ao + (size_t)b * N * C, C, nullptr), 0));
    }
    GemmEpilogue e = vepi(out, C, v->attn_out.bias);
    e.res = x; e.res_dtype = DT_F16; e.ldr = C;
    RCV(r.gemm(B, H, W, C, {GemmSrc{v->ao, C, C, 1}}, v->attn_out.w, C, e, 1));
    return 0;
}

}  // namespace ltt

extern "C" {

int ltt_vae_create(const ltt_vae_config* cfg, int device, ltt_vae** out) {
    if (!cfg || !out) {
        set_error("ltt_vae_create: null argument");
        return -1;
    }
    bool ok = cfg->n_levels >= 1 && cfg->n_levels <= 8 && cfg->out_ch >= 1 && cfg->out_ch <= 4 && cfg->z_channels <= 8 &&
              cfg->embed_dim <= 8 && cfg->num_res_blocks >= 0 && cfg->scale_factor != 0.0f;
    for (int i = 0; ok && i < cfg->n_levels; ++i) ok = cfg->ch * cfg->ch_mult[i] > 0 && (cfg->ch * cfg->ch_mult[i]) % 64 == 0;
    if (!ok) {
        set_error("ltt_vae_create: unsupported decoder configuration (channels must be multiples of 64, out_ch <= 4)");
        return -1;
    }
    LTT_CUDA_OK(cudaSetDevice(device));
    ltt_vae* v = new ltt_vae();
    v->cfg = *cfg;
    v->device = device;
    LTT_CUDA_OK(cudaDeviceGetAttribute(&v->sms, cudaDevAttrMultiProcessorCount, device));
    *out = v;
    return 0;
}

void ltt_vae_destroy(ltt_vae* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaDeviceSynchronize();
    vrelease(v->wptrs);
    vrelease(v->cptrs);
    for (auto& kv : v->params) cudaFree(kv.second.dev);
    delete v;
}

int ltt_vae_load_param(ltt_vae* v, const char* key, const float* data, const int64_t* shape, int ndim, int is_host) {
    if (!v || !key || !data) {
        set_error("ltt_vae_load_param: null argument");
        return -1;
    }
    LTT_CUDA_OK(cudaSetDevice(v->device));
    VParam& p = v->params[key];
    size_t n = 1;
    p.shape.assign(shape, shape + ndim);
    for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
    if (p.dev && p.numel != n) {
        cudaFree(p.dev);
        p.dev = nullptr;
    }
    if (!p.dev) LTT_CUDA_OK(cudaMalloc(&p.dev, n * sizeof(float)));
    p.numel = n;
    LTT_CUDA_OK(cudaMemcpy(p.dev, data, n * sizeof(float), is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
    v->finalized = false;
    return 0;
}

int ltt_vae_finalize(ltt_vae* v) {
    if (!v) return -1;
    if (v->finalized) return 0;             // nothing was loaded since the last call
    LTT_CUDA_OK(cudaSetDevice(v->device));
    RCV(vae_build(v));
    v->finalized = true;
    return 0;
}

int ltt_vae_decode(ltt_vae* v, const float* z, int B, int h, int w, float* img_f32, uint8_t* img_u8, void* stream) {
    if (!v || !v->finalized) {
        set_error("ltt_vae_decode: call ltt_vae_finalize first");
        return -8;
    }
    if (!z || B < 1 || h < 1 || w < 1 || (!img_f32 && !img_u8)) {
        set_error("ltt_vae_decode: bad arguments");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LTT_CUDA_OK(cudaSetDevice(v->device));
    const ltt_vae_config& c = v->cfg;
    if (B != v->B || h != v->h || w != v->w) {
        LTT_CUDA_OK(cudaDeviceSynchronize());
        RCV(vae_workspace(v, B, h, w));
    }
    VRun r{v, st};
    const int N = h * w;
    {
        const int blocks = (int)std::min<size_t>(((size_t)B * N + 255) / 256, 148 * 8);
        LTT_CUDA_OK(launch_k(vae_prep_kernel, dim3(blocks), dim3(256), 0, st, z, v->pq_w, v->pq_b, 1.0f / c.scale_factor, B,
                             c.embed_dim, c.z_channels, N, v->zq));
        v->launches++;
    }
    RCV(conv_in_launch(v->zq, v->conv_in_w, v->conv_in_b, B, c.z_channels, h, w, v->Cmid, v->act0, st));
    v->launches++;
    __half* cur = v->act0;
    auto other = [&](__half* p) { return p == v->act0 ? v->act1 : v->act0; };
    int H = h, W = w;
    RCV(vae_res(r, v->res[0], cur, B, H, W, other(cur))); cur = other(cur);
    RCV(vae_attn(r, cur, B, H, W, other(cur))); cur = other(cur);
    RCV(vae_res(r, v->res[1], cur, B, H, W, other(cur))); cur = other(cur);
    size_t li = 0, ui = 0;
    int ch = v->Cmid;
    for (int lvl = c.n_levels - 1; lvl >= 0; --lvl, ++li) {
        for (int ri : v->level_res[li]) {
            RCV(vae_res(r, v->res[ri], cur, B, H, W, other(cur)));
            cur = other(cur);
            ch = v->res[ri].cout;
        }
        if (lvl != 0) {     // Upsample: nearest x2, then conv3x3 (model.py:55-59)
            const VUp& u = v->ups[ui++];
            RCV(upsample2x_launch(cur, v->tnorm, B, H, W, ch, st));
            v->launches++;
            H *= 2; W *= 2;
            RCV(r.gemm(B, H, W, ch, {GemmSrc{v->tnorm, ch, ch, 9}}, u.conv.w, u.conv.K, vepi(other(cur), ch, u.conv.bias), 1));
            cur = other(cur);
        }
    }
    RCV(r.gn(cur, ch, B, H * W, v->norm_out, 2, v->tnorm));
    {
        const size_t smem = (size_t)c.out_ch * 9 * ch * 2;
        if (smem > 96 * 1024 || (ch & 7)) {
            set_error("vae: conv_out with %d input channels is not supported", ch);
            return -1;
        }
        static bool configured = false;
        if (!configured) {
            LTT_CUDA_OK(cudaFuncSetAttribute(vae_conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = true;
        }
        const size_t total = (size_t)B * H * W, per_block = VO_WARPS * VO_PIX;
        LTT_CUDA_OK(launch_k(vae_conv_out_kernel, dim3((unsigned)((total + per_block - 1) / per_block)), dim3(VO_WARPS * 32), smem, st,
                             v->tnorm, v->out_w, v->out_b, B, H, W, ch, c.out_ch, img_f32, img_u8));
        v->launches++;
    }
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t ltt_vae_launch_count(const ltt_vae* v) { return v ? v->launches : 0; }

}  // extern "C"
