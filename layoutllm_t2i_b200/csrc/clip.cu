// CLIP towers on the sm_100a kernels: the text tower = conditioning prep of the sampler; the vision
// tower + reward head = the CLIP part of train_rl's reward (row f4, /root/reference/models/policy.py:106-123,139).
//
// Mirrors the third-party module the reference calls -- transformers CLIPTextTransformer (token + position embeddings,
// pre-LayerNorm blocks with causal self-attention and a quick-GELU MLP, final LayerNorm, pooled = row of the end-of-text
// token) -- at the reference's call sites: FrozenCLIPEmbedder.forward / encode_one_token
// (/root/reference/GLIGEN/ldm/modules/encoders/modules.py:157-182), get_clip_feature and extract_text_feat
// (/root/reference/txt2img.py:147-156,454-457).  The reference runs one eager fp32 call PER STRING (and, in
// get_clip_feature, a dummy 224 x 224 vision pass with each); here every string of an image -- prompt, negative prompt, up
// to 30 box phrases, the relation phrases -- is a row of ONE batch: 7 launches per layer on [B * L, hidden] rows.
//
// Data flow per layer (x: fp32 residual stream, everything else fp16 with fp32 accumulation):
//   LayerNorm(x) -> h16 | tcgen05 GEMM h16 . Wqkv^T + b -> q, k (head-major rows), v^T | tcgen05 flash attention, causal ->
//   a16 | GEMM a16 . Wo^T + b + x -> x | LayerNorm(x) -> h16 | GEMM h16 . W1^T + b, quick-GELU -> f16 | GEMM f16 . W2^T + b + x -> x
// The reference computes in fp32; fp16 operands bound the error at ~1e-3 relative (tests/clip_checks.py states the gate),
// the same order as the fp16 rounding the UNet's autocast applies to this context at its first use.
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

#define RCC(expr)               \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)

// x[m, :] = token_embedding[ids[m]] + position_embedding[m % L]      (CLIPTextEmbeddings.forward), fp32
__global__ void __launch_bounds__(256) clip_embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok,
                                                         const float* __restrict__ pos, int M, int L, int W, int vocab,
                                                         float* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    const int vec = W >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)M * vec; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / vec), c = (int)(i - (size_t)m * vec);
        int id = ids[m];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);      // ids are validated on the host side of the C-ABI's caller
        const float4 a = reinterpret_cast<const float4*>(tok + (size_t)id * W)[c];
        const float4 b = reinterpret_cast<const float4*>(pos + (size_t)(m % L) * W)[c];
        reinterpret_cast<float4*>(x + (size_t)m * W)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

// y[o] = sum_k x[k] * w[o, k] (+ bias[o]) for o in [o0, o1): one warp per 4 output rows at a time, 16-byte loads, every load of
// a group independent of the others (a one-output-at-a-time loop exposes the full L2 latency per 128 bytes: 389 us for the
// 768 -> 1024 layer of the reward head).  x in shared memory, n % 4 == 0, rows of w 16-byte aligned.
__device__ __forceinline__ void warp_matvec4(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                             int n, int o0, int o1, float* __restrict__ y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int o = o0 + warp * 4; o < o1; o += nwarps * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < n4; k += 32) {
            const float4 xv = x4[k];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (o + j < o1) {
                    const float4 wv = reinterpret_cast<const float4*>(w + (size_t)(o + j) * n)[k];
                    acc[j] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[j]))));
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int s = 16; s; s >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], s);
            if (lane == 0 && o + j < o1) y[o + j] = acc[j] + (bias ? bias[o + j] : 0.f);
        }
    }
}

// pooled[b, :] = hidden[b, e_b, :] with e_b = argmax(ids[b]) (first maximum; eos_token_id == 2) or the first position whose
// id == eos_token_id; text_embeds[b, p] = sum_k pooled[b, k] * proj[p, k] in fp32 (CLIPModel.text_projection, no bias).
// One CTA per sequence.
__global__ void __launch_bounds__(256) clip_pool_kernel(const int32_t* __restrict__ ids, const float* __restrict__ hidden, int L,
                                                        int W, int eos_token_id, const float* __restrict__ proj, int P,
                                                        float* __restrict__ pooled, float* __restrict__ embeds) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float prow[];
    __shared__ int epos;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int best = 0;
        if (eos_token_id == 2) {
            int mx = ids[(size_t)b * L];
            for (int i = 1; i < L; ++i) {
                const int v = ids[(size_t)b * L + i];
                if (v > mx) { mx = v; best = i; }
            }
        } else {
            for (int i = 0; i < L; ++i)
                if (ids[(size_t)b * L + i] == eos_token_id) { best = i; break; }
        }
        epos = best;
    }
    __syncthreads();
    const float* src = hidden + ((size_t)b * L + epos) * W;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const float v = src[i];
        prow[i] = v;
        if (pooled && blockIdx.y == 0) pooled[(size_t)b * W + i] = v;
    }
    __syncthreads();
    if (!embeds) return;
    // the projection rows are split over gridDim.y CTAs (each re-gathers the 3 KB row)
    const int per = ((P + (int)gridDim.y - 1) / (int)gridDim.y + 3) & ~3;
    const int p0 = (int)blockIdx.y * per, p1 = min(P, p0 + per);
    warp_matvec4(prow, proj, nullptr, W, p0, p1, embeds + (size_t)b * P);
}

// ---- vision tower (transformers CLIPVisionEmbeddings / CLIPVisionTransformer)
// patches[(b * G * G + py * G + px), c * P * P + ky * P + kx] = fp16(pixel[b, c, py * P + ky, px * P + kx]), zero padded to Kpad:
// the stride-P patch convolution becomes one GEMM against the flattened [W, 3 P P] kernel
__global__ void __launch_bounds__(256) clipv_patches_kernel(const float* __restrict__ px, int B, int S, int P, int G, int Kpad,
                                                            __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int K = 3 * P * P;
    const size_t total = (size_t)B * G * G * Kpad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad);
        const size_t row = i / Kpad;
        float v = 0.f;
        if (k < K) {
            const int b = (int)(row / (G * G)), pr = (int)(row % (G * G)), py = pr / G, pxx = pr % G;
            const int c = k / (P * P), kk = k % (P * P), ky = kk / P, kx = kk % P;
            v = px[(((size_t)b * 3 + c) * S + py * P + ky) * S + pxx * P + kx];
        }
        out[i] = __float2half_rn(v);
    }
}
// [W, 3, P, P] fp32 -> [W, Kpad] fp16 (zero padded rows)
__global__ void clipv_pack_patch_kernel(const float* __restrict__ w, int W, int K, int Kpad, __half* __restrict__ out) {
    const size_t total = (size_t)W * Kpad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % Kpad);
        out[i] = __float2half_rn(k < K ? w[(i / Kpad) * K + k] : 0.f);
    }
}
// x[b, 0, :] = class_embedding + pos[0];  x[b, 1 + p, :] = patch_out[b * NP + p, :] + pos[1 + p]      (fp32)
__global__ void __launch_bounds__(256) clipv_assemble_kernel(const float* __restrict__ pe, const float* __restrict__ cls,
                                                             const float* __restrict__ pos, int B, int T, int W,
                                                             float* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    const int vec = W >> 2;
    const size_t total = (size_t)B * T * vec;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % vec);
        const size_t row = i / vec;
        const int b = (int)(row / T), t = (int)(row % T);
        const float4 a = t == 0 ? reinterpret_cast<const float4*>(cls)[c]
                                : reinterpret_cast<const float4*>(pe + ((size_t)b * (T - 1) + t - 1) * W)[c];
        const float4 p = reinterpret_cast<const float4*>(pos + (size_t)t * W)[c];
        reinterpret_cast<float4*>(x + row * W)[c] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    return s;
}
// pooled[b] = post_layernorm(hidden[b, 0, :]); embeds[b, p] = sum_k pooled[b, k] proj[p, k]   (fp32; one CTA per image)
__global__ void __launch_bounds__(256) clipv_pool_kernel(const float* __restrict__ hidden, int T, int W, const float* __restrict__ g,
                                                         const float* __restrict__ bt, float eps, const float* __restrict__ proj,
                                                         int P, float* __restrict__ pooled, float* __restrict__ embeds) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float prow[];
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float* src = hidden + (size_t)b * T * W;
    float s = 0.f;
    for (int i = threadIdx.x; i < W; i += 256) {
        prow[i] = src[i];
        s += prow[i];
    }
    const float mean = block_sum_256(s, red) / W;
    float ss = 0.f;
    for (int i = threadIdx.x; i < W; i += 256) {
        const float d = prow[i] - mean;
        ss += d * d;
    }
    const float rstd = rsqrtf(block_sum_256(ss, red) / W + eps);
    for (int i = threadIdx.x; i < W; i += 256) {
        const float y = (prow[i] - mean) * rstd * g[i] + bt[i];
        prow[i] = y;
        if (pooled && blockIdx.y == 0) pooled[(size_t)b * W + i] = y;
    }
    __syncthreads();
    if (!embeds) return;
    const int per = ((P + (int)gridDim.y - 1) / (int)gridDim.y + 3) & ~3;
    const int p0 = (int)blockIdx.y * per, p1 = min(P, p0 + per);
    warp_matvec4(prow, proj, nullptr, W, p0, p1, embeds + (size_t)b * P);
}

// ---- image preprocessing in front of the vision tower (transformers CLIPImageProcessor with the PIL backend: Pillow's
// two-pass BICUBIC resize -- src/libImaging/Resample.c, 8-bit rounding after each pass, horizontal first -- centre crop,
// rescale, normalise).  Integer arithmetic on bytes: bit-exact against Pillow.
constexpr int PRE_BITS = 32 - 8 - 2;      // Resample.c PRECISION_BITS
__device__ __forceinline__ uint8_t pre_clip8(int v) {
    v >>= PRE_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// in [B, H, W, 3] u8 -> tmp [B, H, rw, 3] u8
__global__ void __launch_bounds__(256) clip_resize_h_kernel(const uint8_t* __restrict__ in, int B, int H, int W, int rw,
                                                            const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                                                            uint8_t* __restrict__ tmp) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * H * rw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % rw);
        const size_t row = i / rw;                  // b * H + y
        const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
        const uint8_t* src = in + (row * W + xmin) * 3;
        const int* k = kk + (size_t)xo * ksize;
        int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
        for (int x = 0; x < n; ++x) {
            const int w = k[x];
            a0 += src[3 * x] * w;
            a1 += src[3 * x + 1] * w;
            a2 += src[3 * x + 2] * w;
        }
        uint8_t* dst = tmp + i * 3;
        dst[0] = pre_clip8(a0); dst[1] = pre_clip8(a1); dst[2] = pre_clip8(a2);
    }
}
// tmp [B, H, rw, 3] u8 -> vertical pass at the cropped positions only -> out [B, 3, S, S] fp32 through the [3][256] table
__global__ void __launch_bounds__(256) clip_resize_v_kernel(const uint8_t* __restrict__ tmp, int B, int H, int rw, int S, int top,
                                                            int left, const int* __restrict__ bounds, const int* __restrict__ kk,
                                                            int ksize, const float* __restrict__ lut, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * S * S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % S), yo = (int)((i / S) % S), b = (int)(i / ((size_t)S * S));
        const int yy = yo + top;
        const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
        const uint8_t* src = tmp + (((size_t)b * H + ymin) * rw + xo + left) * 3;
        const int* k = kk + (size_t)yy * ksize;
        int a0 = 1 << (PRE_BITS - 1), a1 = a0, a2 = a0;
        for (int y = 0; y < n; ++y) {
            const int w = k[y];
            const uint8_t* p = src + (size_t)y * rw * 3;
            a0 += p[0] * w;
            a1 += p[1] * w;
            a2 += p[2] * w;
        }
        const size_t plane = (size_t)S * S, o = (size_t)b * 3 * plane + (size_t)yo * S + xo;
        out[o] = lut[pre_clip8(a0)];
        out[o + plane] = lut[256 + pre_clip8(a1)];
        out[o + 2 * plane] = lut[512 + pre_clip8(a2)];
    }
}

// ---- reward head (Reward.forward, models/policy.py:115-139; AestheticMLP tools/aesthetic.py:15-31): one CTA per sample.
// clip = cos(txt, pred) + cos(gt, pred); aes = the five bias-affine layers (eval: dropout is the identity, there is no
// activation) on pred / |pred|; reward = clip + 0.1 aes + 10 miou + 10 laysim.   D <= 1024, layer widths <= 1024.
struct AesW {
    const float* w[5];
    const float* b[5];
    int dim[6];
};
// Launched as one cluster of RH_CLUSTER CTAs per sample: every CTA normalises the (3 KB) feature rows itself, the 768 -> 1024
// first layer -- 90 % of the arithmetic -- is split over the cluster, rank 0 gathers the slices through distributed shared
// memory and finishes the (small) remaining layers.
constexpr int RH_CLUSTER = 8;
__global__ void __launch_bounds__(256) reward_head_kernel(const float* __restrict__ txt, const float* __restrict__ pred,
                                                          const float* __restrict__ gt, int D, AesW aw,
                                                          const float* __restrict__ miou, const float* __restrict__ laysim,
                                                          float* __restrict__ reward, float* __restrict__ clip_out,
                                                          float* __restrict__ aes_out) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float buf[2][1024];
    __shared__ float red[8];
    const int b = blockIdx.y;
    const uint32_t rank = cluster_ctarank();
    const float *t = txt + (size_t)b * D, *p = pred + (size_t)b * D, *g = gt + (size_t)b * D;
    float tt = 0.f, pp = 0.f, gg = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) {
        tt += t[i] * t[i];
        pp += p[i] * p[i];
        gg += g[i] * g[i];
    }
    // F.normalize: x / max(|x|, 1e-12)
    const float nt = fmaxf(sqrtf(block_sum_256(tt, red)), 1e-12f), np_ = fmaxf(sqrtf(block_sum_256(pp, red)), 1e-12f),
                ng = fmaxf(sqrtf(block_sum_256(gg, red)), 1e-12f);
    float tp = 0.f, gp = 0.f, p2 = 0.f;
    for (int i = threadIdx.x; i < D; i += 256) {
        const float pn = p[i] / np_;
        buf[0][i] = pn;
        tp += (t[i] / nt) * pn;
        gp += (g[i] / ng) * pn;
        p2 += pn * pn;
    }
    const float clip = block_sum_256(tp, red) + block_sum_256(gp, red);
    float n2 = sqrtf(block_sum_256(p2, red));       // `normalized`: divide by the norm of the (already normalised) row, 0 -> 1
    if (n2 == 0.f) n2 = 1.f;
    for (int i = threadIdx.x; i < D; i += 256) buf[0][i] /= n2;
    __syncthreads();
    // layer 0: this CTA's slice of the outputs, written at its final position of buf[1]
    const int out0 = aw.dim[1];
    const int per = ((out0 + RH_CLUSTER - 1) / RH_CLUSTER + 3) & ~3;
    const int o0 = min(out0, (int)rank * per), o1 = min(out0, o0 + per);
    warp_matvec4(buf[0], aw.w[0], aw.b[0], aw.dim[0], o0, o1, buf[1]);
    cluster_sync_all();
    if (rank == 0) {
        const uint32_t base = smem_u32(&buf[1][0]);
        for (int i = per + threadIdx.x; i < out0; i += 256) buf[1][i] = dsmem_ld_f32(dsmem_map(base + 4u * i, (uint32_t)(i / per)));
    }
    cluster_sync_all();          // no CTA leaves while rank 0 may still read its slice
    if (rank != 0) return;
    int cur = 1;
#pragma unroll 1
    for (int l = 1; l < 5; ++l) {
        warp_matvec4(buf[cur], aw.w[l], aw.b[l], aw.dim[l], 0, aw.dim[l + 1], buf[cur ^ 1]);
        __syncthreads();
        cur ^= 1;
    }
    if (threadIdx.x == 0) {
        const float aes = buf[cur][0];
        if (clip_out) clip_out[b] = clip;
        if (aes_out) aes_out[b] = aes;
        reward[b] = clip + aes * 0.1f + (miou ? miou[b] * 10.f : 0.f) + (laysim ? laysim[b] * 10.f : 0.f);
    }
}

struct CParam { float* dev = nullptr; std::vector<int64_t> shape; size_t numel = 0; };
typedef std::map<std::string, CParam> CParams;
struct CLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    __half *wqkv, *wo, *w1, *w2;
    const float *bqkv, *bo, *b1, *b2;
};
// the encoder both towers share (transformers CLIPEncoder): weights + the activation workspace of one (B, L)
struct Tower {
    int W = 0, heads = 0, ffn = 0;
    float eps = 1e-5f;
    std::vector<CLayer> layers;
    std::vector<void*> wptrs, cptrs;
    int B = 0, L = 0, pitch_v = 0;
    float* x = nullptr;          // fp32 residual stream [B * L, W]
    __half *h16 = nullptr, *q = nullptr, *k = nullptr, *vt = nullptr, *a16 = nullptr, *f16 = nullptr;
};

static int calloc_dev(std::vector<void*>& arena, void** out, size_t bytes) {
    void* p = nullptr;
    LTT_CUDA_OK(cudaMalloc(&p, bytes ? bytes : 16));
    arena.push_back(p);
    *out = p;
    return 0;
}
static void crelease(std::vector<void*>& arena) {
    for (void* p : arena) cudaFree(p);
    arena.clear();
}
static const CParam* cfind(const CParams& params, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = params.find(key);
    if (it == params.end()) {
        set_error("clip: missing parameter '%s' (load_state_dict incomplete)", key.c_str());
        return nullptr;
    }
    if (it->second.shape != std::vector<int64_t>(shape)) {
        set_error("clip: parameter '%s' has the wrong shape for this configuration", key.c_str());
        return nullptr;
    }
    return &it->second;
}
#define CGET(var, key, ...)                                   \
    const CParam* var = cfind(params, (key), {__VA_ARGS__});  \
    if (!var) return -6;

static int load_param(CParams& params, int device, const char* key, const float* data, const int64_t* shape, int ndim, int is_host) {
    LTT_CUDA_OK(cudaSetDevice(device));
    CParam& p = params[key];
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
    return 0;
}

// encoder.layers.* of one tower (`prefix` = "text_model." / "vision_model.") -> packed fp16 weights
// `packed` collects the keys of the fp32 matrices that live on as packed fp16 copies (released by the caller after the packing
// kernels have run: biases, norms, embeddings and projections stay fp32)
static int tower_build(Tower& t, const CParams& params, const std::string& prefix, int n_layers, std::vector<std::string>& packed) {
    const int64_t W = t.W, F = t.ffn;
    crelease(t.wptrs);
    t.layers.clear();
    for (int i = 0; i < n_layers; ++i) {
        const std::string p = prefix + "encoder.layers." + std::to_string(i) + ".";
        CLayer l{};
        CGET(g1, p + "layer_norm1.weight", W) CGET(b1, p + "layer_norm1.bias", W)
        CGET(g2, p + "layer_norm2.weight", W) CGET(b2, p + "layer_norm2.bias", W)
        l.ln1_g = g1->dev; l.ln1_b = b1->dev; l.ln2_g = g2->dev; l.ln2_b = b2->dev;
        void *wq, *bq;
        RCC(calloc_dev(t.wptrs, &wq, (size_t)3 * W * W * 2));
        RCC(calloc_dev(t.wptrs, &bq, (size_t)3 * W * 4));
        const char* names[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; ++j) {       // q / k / v stacked into one [3W, W] matrix (the softmax scale stays in the attention kernel)
            CGET(w, p + "self_attn." + names[j] + ".weight", W, W)
            CGET(b, p + "self_attn." + names[j] + ".bias", W)
            packed.push_back(p + "self_attn." + names[j] + ".weight");
            RCC(pack_rows_launch(w->dev, (int)W, (int)W, (__half*)wq, j * (int)W, 0, 0));
            LTT_CUDA_OK(cudaMemcpy((float*)bq + j * W, b->dev, W * 4, cudaMemcpyDeviceToDevice));
        }
        l.wqkv = (__half*)wq; l.bqkv = (const float*)bq;
        auto lin = [&](const std::string& name, int64_t N, int64_t K, __half** wout, const float** bout) -> int {
            CGET(w, p + name + ".weight", N, K)
            CGET(b, p + name + ".bias", N)
            packed.push_back(p + name + ".weight");
            void* q;
            RCC(calloc_dev(t.wptrs, &q, (size_t)N * K * 2));
            RCC(pack_rows_launch(w->dev, (int)N, (int)K, (__half*)q, 0, 0, 0));
            *wout = (__half*)q; *bout = b->dev;
            return 0;
        };
        RCC(lin("self_attn.out_proj", W, W, &l.wo, &l.bo));
        RCC(lin("mlp.fc1", F, W, &l.w1, &l.b1));
        RCC(lin("mlp.fc2", W, F, &l.w2, &l.b2));
        t.layers.push_back(l);
    }
    return 0;
}

static void release_packed(CParams& params, const std::vector<std::string>& packed) {
    for (const std::string& k : packed) {
        auto it = params.find(k);
        if (it != params.end()) {
            cudaFree(it->second.dev);
            params.erase(it);
        }
    }
}

static int tower_workspace(Tower& t, int B, int L) {
    crelease(t.cptrs);
    const size_t M = (size_t)B * L, W = t.W;
    t.pitch_v = (L + 7) & ~7;
    auto A = [&](auto** p, size_t bytes) { return calloc_dev(t.cptrs, (void**)p, bytes); };
    RCC(A(&t.x, M * W * 4));
    RCC(A(&t.h16, M * W * 2)); RCC(A(&t.q, M * W * 2)); RCC(A(&t.k, M * W * 2)); RCC(A(&t.a16, M * W * 2));
    RCC(A(&t.vt, (size_t)B * W * t.pitch_v * 2));
    RCC(A(&t.f16, M * (size_t)t.ffn * 2));
    LTT_CUDA_OK(cudaMemset(t.vt, 0, (size_t)B * W * t.pitch_v * 2));      // pitch padding columns are never written
    t.B = B; t.L = L;
    return 0;
}

// rows = nb sequences x len tokens; the QKV projection passes the real geometry (its epilogue transposes V per sequence),
// the other GEMMs run on the flat [B * L] row list (nb = 1)
static int clip_gemm(int sms, cudaStream_t st, int nb, int len, int N, const __half* a, int K, const __half* w, const GemmEpilogue& epi) {
    GemmProblem p{};
    p.B = nb; p.H = 1; p.W = len; p.N = N; p.nsrc = 1;
    p.src[0] = GemmSrc{a, K, K, 1};
    p.w = w; p.Ktot = K; p.w_static = 1; p.epi = epi;
    return gemm_tc_launch(p, sms, st);
}

// CLIPEncoder.forward on t.x in place: 7 launches per layer
static int tower_run(Tower& t, int sms, cudaStream_t st, int causal, int64_t& launches) {
    const int B = t.B, L = t.L, M = B * L, W = t.W;
    for (const CLayer& l : t.layers) {
        RCC(layernorm_launch(t.x, DT_F32, M, W, l.ln1_g, l.ln1_b, t.eps, t.h16, nullptr, st));
        {
            GemmEpilogue e;
            e.bias = l.bqkv;
            e.out_mode = OUT_QKV;
            e.q = t.q; e.k = t.k; e.vt = t.vt;
            e.C = W; e.dhead = 64; e.dpad = 64; e.rows_q = L; e.rows_k = L; e.pitch_v = t.pitch_v; e.tokens = L; e.qkv_base = 0;
            RCC(clip_gemm(sms, st, B, L, 3 * W, t.h16, W, l.wqkv, e));
        }
        {
            AttnProblem p{};
            p.B = B; p.heads = t.heads; p.dhead = 64; p.dpad = 64; p.nq = L; p.nk = L;
            p.q = t.q; p.rows_q = L; p.k = t.k; p.rows_k = L; p.vt = t.vt; p.pitch_v = t.pitch_v;
            p.out = t.a16; p.ldo = W; p.scale = 0.125f; p.causal = causal;
            RCC(attn_tc_launch(p, st));
        }
        {
            GemmEpilogue e;
            e.bias = l.bo; e.res = t.x; e.res_dtype = DT_F32; e.ldr = W; e.out = t.x; e.out_dtype = DT_F32; e.ldo = W;
            RCC(clip_gemm(sms, st, 1, M, W, t.a16, W, l.wo, e));
        }
        RCC(layernorm_launch(t.x, DT_F32, M, W, l.ln2_g, l.ln2_b, t.eps, t.h16, nullptr, st));
        {
            GemmEpilogue e;
            e.bias = l.b1; e.act = ACT_QUICKGELU; e.out = t.f16; e.out_dtype = DT_F16; e.ldo = t.ffn;
            RCC(clip_gemm(sms, st, 1, M, t.ffn, t.h16, W, l.w1, e));
        }
        {
            GemmEpilogue e;
            e.bias = l.b2; e.res = t.x; e.res_dtype = DT_F32; e.ldr = W; e.out = t.x; e.out_dtype = DT_F32; e.ldo = W;
            RCC(clip_gemm(sms, st, 1, M, W, t.f16, t.ffn, l.w2, e));
        }
        launches += 7;
    }
    return 0;
}

}  // namespace ltt

using namespace ltt;

struct ltt_clip {
    ltt_clip_config cfg;
    int device = 0, sms = 148;
    CParams params;
    bool finalized = false;
    Tower t;
    const float *tok = nullptr, *pos = nullptr, *fin_g = nullptr, *fin_b = nullptr, *proj = nullptr;
    float* hid = nullptr;
    std::vector<void*> hptrs;
    int64_t launches = 0;
};

struct ltt_clip_vision {
    ltt_clip_vision_config cfg;
    int device = 0, sms = 148;
    CParams params;
    bool finalized = false;
    Tower t;
    int G = 0, NP = 0, Kpad = 0;           // patches per side, patches per image, padded 3 P P
    const float *cls = nullptr, *pos = nullptr, *pre_g = nullptr, *pre_b = nullptr, *post_g = nullptr, *post_b = nullptr, *proj = nullptr;
    __half* patch_w = nullptr;
    std::vector<void*> wptrs, cptrs;
    int B = 0;
    __half* patches = nullptr;
    float *pe = nullptr, *xa = nullptr;
    // preprocessing workspace for (B, H, W) + normalisation table
    std::vector<void*> pptrs;
    int pB = 0, pH = 0, pW = 0, p_rw = 0, p_rh = 0, p_top = 0, p_left = 0, p_ksx = 0, p_ksy = 0;
    float p_mean[3] = {0, 0, 0}, p_std[3] = {0, 0, 0};
    int *p_bx = nullptr, *p_kx = nullptr, *p_by = nullptr, *p_ky = nullptr;
    float* p_lut = nullptr;
    uint8_t* p_tmp = nullptr;
    int64_t launches = 0;
};

namespace ltt {

// Resample.c bicubic_filter / precompute_coeffs / normalize_coeffs_8bpc for the full-image box, in the same double arithmetic
static double pil_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static int pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    double scale = (double)in_size / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    kk.assign((size_t)out_size * ksize, 0);
    const double ss = 1.0 / filterscale;
    std::vector<double> w(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = pil_bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[x] / ww : w[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRE_BITS)) : (int)(0.5 + v * (1 << PRE_BITS));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

}  // namespace ltt

extern "C" {

// ---------------------------------------------------------------------------------------------------- text tower
int ltt_clip_create(const ltt_clip_config* cfg, int device, ltt_clip** out) {
    if (!cfg || !out) {
        set_error("ltt_clip_create: null argument");
        return -1;
    }
    const bool ok = cfg->vocab >= 1 && cfg->max_pos >= 1 && cfg->hidden >= 64 && cfg->heads >= 1 && cfg->hidden == cfg->heads * 64 &&
                    cfg->ffn >= 64 && cfg->ffn % 64 == 0 && cfg->layers >= 1 && cfg->act == 0 && cfg->proj_dim >= 0 && cfg->eps > 0.f;
    if (!ok) {
        set_error("ltt_clip_create: unsupported text tower (heads of 64, hidden / ffn multiples of 64, quick_gelu)");
        return -1;
    }
    LTT_CUDA_OK(cudaSetDevice(device));
    ltt_clip* c = new ltt_clip();
    c->cfg = *cfg;
    c->device = device;
    c->t.W = cfg->hidden; c->t.heads = cfg->heads; c->t.ffn = cfg->ffn; c->t.eps = cfg->eps;
    LTT_CUDA_OK(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    *out = c;
    return 0;
}

void ltt_clip_destroy(ltt_clip* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    crelease(c->t.wptrs);
    crelease(c->t.cptrs);
    crelease(c->hptrs);
    for (auto& kv : c->params) cudaFree(kv.second.dev);
    delete c;
}

int ltt_clip_load_param(ltt_clip* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_host) {
    if (!c || !key || !data || (ndim > 0 && !shape)) {
        set_error("ltt_clip_load_param: null argument");
        return -1;
    }
    c->finalized = false;
    return load_param(c->params, c->device, key, data, shape, ndim, is_host);
}

int ltt_clip_finalize(ltt_clip* c) {
    if (!c) return -1;
    if (c->finalized) return 0;             // nothing was loaded since the last call
    LTT_CUDA_OK(cudaSetDevice(c->device));
    const ltt_clip_config& g = c->cfg;
    const CParams& params = c->params;
    const int64_t W = g.hidden;
    const std::string tm = "text_model.";
    CGET(t, tm + "embeddings.token_embedding.weight", g.vocab, W)
    CGET(p, tm + "embeddings.position_embedding.weight", g.max_pos, W)
    CGET(fg, tm + "final_layer_norm.weight", W)
    CGET(fb, tm + "final_layer_norm.bias", W)
    c->tok = t->dev; c->pos = p->dev; c->fin_g = fg->dev; c->fin_b = fb->dev;
    c->proj = nullptr;
    if (g.proj_dim > 0) {
        CGET(pw, "text_projection.weight", g.proj_dim, W)
        c->proj = pw->dev;
    }
    std::vector<std::string> packed;
    RCC(tower_build(c->t, params, tm, g.layers, packed));
    LTT_CUDA_OK(cudaDeviceSynchronize());
    release_packed(c->params, packed);      // 2/3 of the tower's device memory; a later finalize needs the state_dict loaded again
    c->finalized = true;
    return 0;
}

int ltt_clip_encode(ltt_clip* c, const int32_t* ids, int B, int L, float* last_hidden, float* pooled, float* text_embeds,
                    void* stream) {
    if (!c || !c->finalized) {
        set_error("ltt_clip_encode: call ltt_clip_finalize first");
        return -8;
    }
    const ltt_clip_config& g = c->cfg;
    if (!ids || B < 1 || L < 1 || L > g.max_pos || (!last_hidden && !pooled && !text_embeds)) {
        set_error("ltt_clip_encode: bad arguments (B=%d, L=%d, max_position_embeddings=%d)", B, L, g.max_pos);
        return -1;
    }
    if (text_embeds && !c->proj) {
        set_error("ltt_clip_encode: text_embeds requested but the tower has no text_projection (proj_dim = 0)");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LTT_CUDA_OK(cudaSetDevice(c->device));
    Tower& t = c->t;
    if (B != t.B || L != t.L) {
        LTT_CUDA_OK(cudaDeviceSynchronize());
        RCC(tower_workspace(t, B, L));
        crelease(c->hptrs);
        RCC(calloc_dev(c->hptrs, (void**)&c->hid, (size_t)B * L * g.hidden * 4));
    }
    const int M = B * L, W = g.hidden;
    {
        const int blocks = (int)std::min<size_t>(((size_t)M * (W >> 2) + 255) / 256, (size_t)c->sms * 8);
        LTT_CUDA_OK(launch_k(clip_embed_kernel, dim3(blocks), dim3(256), 0, st, ids, c->tok, c->pos, M, L, W, g.vocab, t.x));
        c->launches++;
    }
    RCC(tower_run(t, c->sms, st, 1, c->launches));
    float* hid = last_hidden ? last_hidden : c->hid;
    RCC(layernorm_launch(t.x, DT_F32, M, W, c->fin_g, c->fin_b, g.eps, nullptr, hid, st));
    c->launches++;
    if (pooled || text_embeds) {
        LTT_CUDA_OK(launch_k(clip_pool_kernel, dim3(B, text_embeds ? 8 : 1), dim3(256), (size_t)W * 4, st, ids, (const float*)hid, L, W, g.eos_token_id,
                             text_embeds ? c->proj : (const float*)nullptr, g.proj_dim, pooled, text_embeds));
        c->launches++;
    }
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t ltt_clip_launch_count(const ltt_clip* c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------------------------------- vision tower
int ltt_clip_vision_create(const ltt_clip_vision_config* cfg, int device, ltt_clip_vision** out) {
    if (!cfg || !out) {
        set_error("ltt_clip_vision_create: null argument");
        return -1;
    }
    const bool ok = cfg->image_size >= cfg->patch && cfg->patch >= 1 && cfg->image_size % cfg->patch == 0 && cfg->hidden >= 64 &&
                    cfg->hidden == cfg->heads * 64 && cfg->hidden <= 4096 && cfg->ffn >= 64 && cfg->ffn % 64 == 0 && cfg->layers >= 1 &&
                    cfg->act == 0 && cfg->proj_dim >= 0 && cfg->eps > 0.f;
    if (!ok) {
        set_error("ltt_clip_vision_create: unsupported vision tower (heads of 64, hidden / ffn multiples of 64, quick_gelu)");
        return -1;
    }
    LTT_CUDA_OK(cudaSetDevice(device));
    ltt_clip_vision* c = new ltt_clip_vision();
    c->cfg = *cfg;
    c->device = device;
    c->G = cfg->image_size / cfg->patch;
    c->NP = c->G * c->G;
    c->Kpad = (3 * cfg->patch * cfg->patch + 63) / 64 * 64;
    c->t.W = cfg->hidden; c->t.heads = cfg->heads; c->t.ffn = cfg->ffn; c->t.eps = cfg->eps;
    LTT_CUDA_OK(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    *out = c;
    return 0;
}

void ltt_clip_vision_destroy(ltt_clip_vision* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    crelease(c->t.wptrs);
    crelease(c->t.cptrs);
    crelease(c->wptrs);
    crelease(c->cptrs);
    crelease(c->pptrs);
    for (auto& kv : c->params) cudaFree(kv.second.dev);
    delete c;
}

int ltt_clip_vision_load_param(ltt_clip_vision* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_host) {
    if (!c || !key || !data || (ndim > 0 && !shape)) {
        set_error("ltt_clip_vision_load_param: null argument");
        return -1;
    }
    c->finalized = false;
    return load_param(c->params, c->device, key, data, shape, ndim, is_host);
}

int ltt_clip_vision_finalize(ltt_clip_vision* c) {
    if (!c) return -1;
    if (c->finalized) return 0;
    LTT_CUDA_OK(cudaSetDevice(c->device));
    const ltt_clip_vision_config& g = c->cfg;
    const CParams& params = c->params;
    const int64_t W = g.hidden, P = g.patch;
    const std::string vm = "vision_model.";
    CGET(cls, vm + "embeddings.class_embedding", W)
    CGET(pw, vm + "embeddings.patch_embedding.weight", W, 3, P, P)
    CGET(pos, vm + "embeddings.position_embedding.weight", c->NP + 1, W)
    CGET(g0, vm + "pre_layrnorm.weight", W) CGET(b0, vm + "pre_layrnorm.bias", W)
    CGET(g1, vm + "post_layernorm.weight", W) CGET(b1, vm + "post_layernorm.bias", W)
    c->cls = cls->dev; c->pos = pos->dev; c->pre_g = g0->dev; c->pre_b = b0->dev; c->post_g = g1->dev; c->post_b = b1->dev;
    c->proj = nullptr;
    if (g.proj_dim > 0) {
        CGET(vp, "visual_projection.weight", g.proj_dim, W)
        c->proj = vp->dev;
    }
    crelease(c->wptrs);
    RCC(calloc_dev(c->wptrs, (void**)&c->patch_w, (size_t)W * c->Kpad * 2));
    clipv_pack_patch_kernel<<<256, 256>>>(pw->dev, (int)W, (int)(3 * P * P), c->Kpad, c->patch_w);
    LTT_CUDA_OK(cudaGetLastError());
    std::vector<std::string> packed;
    RCC(tower_build(c->t, params, vm, g.layers, packed));
    LTT_CUDA_OK(cudaDeviceSynchronize());
    release_packed(c->params, packed);
    c->finalized = true;
    return 0;
}

int ltt_clip_vision_encode(ltt_clip_vision* c, const float* pixel_values, int B, float* last_hidden, float* pooled, float* image_embeds,
                           void* stream) {
    if (!c || !c->finalized) {
        set_error("ltt_clip_vision_encode: call ltt_clip_vision_finalize first");
        return -8;
    }
    const ltt_clip_vision_config& g = c->cfg;
    if (!pixel_values || B < 1 || (!last_hidden && !pooled && !image_embeds)) {
        set_error("ltt_clip_vision_encode: bad arguments");
        return -1;
    }
    if (image_embeds && !c->proj) {
        set_error("ltt_clip_vision_encode: image_embeds requested but the tower has no visual_projection (proj_dim = 0)");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    LTT_CUDA_OK(cudaSetDevice(c->device));
    Tower& t = c->t;
    const int T = c->NP + 1, W = g.hidden, M = B * T;
    if (B != c->B) {
        LTT_CUDA_OK(cudaDeviceSynchronize());
        RCC(tower_workspace(t, B, T));
        crelease(c->cptrs);
        RCC(calloc_dev(c->cptrs, (void**)&c->patches, (size_t)B * c->NP * c->Kpad * 2));
        RCC(calloc_dev(c->cptrs, (void**)&c->pe, (size_t)B * c->NP * W * 4));
        RCC(calloc_dev(c->cptrs, (void**)&c->xa, (size_t)M * W * 4));
        c->B = B;
    }
    {
        const size_t n = (size_t)B * c->NP * c->Kpad;
        LTT_CUDA_OK(launch_k(clipv_patches_kernel, dim3((unsigned)std::min<size_t>((n + 255) / 256, (size_t)c->sms * 16)), dim3(256), 0, st,
                             pixel_values, B, g.image_size, g.patch, c->G, c->Kpad, c->patches));
        GemmEpilogue e;
        e.out = c->pe; e.out_dtype = DT_F32; e.ldo = W;
        RCC(clip_gemm(c->sms, st, 1, B * c->NP, W, c->patches, c->Kpad, c->patch_w, e));
        const size_t m = (size_t)M * (W >> 2);
        LTT_CUDA_OK(launch_k(clipv_assemble_kernel, dim3((unsigned)std::min<size_t>((m + 255) / 256, (size_t)c->sms * 16)), dim3(256), 0, st,
                             (const float*)c->pe, c->cls, c->pos, B, T, W, c->xa));
        RCC(layernorm_launch(c->xa, DT_F32, M, W, c->pre_g, c->pre_b, g.eps, nullptr, t.x, st));
        c->launches += 4;
    }
    RCC(tower_run(t, c->sms, st, 0, c->launches));
    if (last_hidden) {
        LTT_CUDA_OK(cudaMemcpyAsync(last_hidden, t.x, (size_t)M * W * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (pooled || image_embeds) {
        LTT_CUDA_OK(launch_k(clipv_pool_kernel, dim3(B, image_embeds ? 8 : 1), dim3(256), (size_t)W * 4, st, (const float*)t.x, T, W, c->post_g, c->post_b, g.eps,
                             image_embeds ? c->proj : (const float*)nullptr, g.proj_dim, pooled, image_embeds));
        c->launches++;
    }
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int ltt_clip_vision_preprocess(ltt_clip_vision* c, const uint8_t* images, int B, int H, int W, const float* mean, const float* std_,
                               float* pixel_values, void* stream) {
    if (!c || !images || !mean || !std_ || !pixel_values || B < 1 || H < 1 || W < 1) {
        set_error("ltt_clip_vision_preprocess: bad arguments");
        return -1;
    }
    const int S = c->cfg.image_size;
    cudaStream_t st = (cudaStream_t)stream;
    LTT_CUDA_OK(cudaSetDevice(c->device));
    const bool same_norm = memcmp(mean, c->p_mean, 12) == 0 && memcmp(std_, c->p_std, 12) == 0;
    if (B != c->pB || H != c->pH || W != c->pW || !same_norm) {
        LTT_CUDA_OK(cudaDeviceSynchronize());
        crelease(c->pptrs);
        // shortest edge -> S, long edge int(S * long / short) (transformers get_resize_output_image_size), centre crop S x S
        const int shortE = W <= H ? W : H, longE = W <= H ? H : W;
        const int new_long = (int)((double)S * longE / shortE);
        c->p_rh = W <= H ? new_long : S;
        c->p_rw = W <= H ? S : new_long;
        c->p_top = (c->p_rh - S) / 2;
        c->p_left = (c->p_rw - S) / 2;
        std::vector<int> bx, kx, by, ky;
        c->p_ksx = pil_coeffs(W, c->p_rw, bx, kx);
        c->p_ksy = pil_coeffs(H, c->p_rh, by, ky);
        std::vector<float> lut(3 * 256);
        for (int ch = 0; ch < 3; ++ch)
            for (int u = 0; u < 256; ++u) {
                const float v = (float)((double)u * (1.0 / 255.0));      // rescale in float64, cast to float32
                lut[ch * 256 + u] = (v - mean[ch]) / std_[ch];          // normalize in float32
            }
        auto up = [&](auto** dst, const void* src, size_t bytes) -> int {
            RCC(calloc_dev(c->pptrs, (void**)dst, bytes));
            LTT_CUDA_OK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
            return 0;
        };
        RCC(up(&c->p_bx, bx.data(), bx.size() * 4)); RCC(up(&c->p_kx, kx.data(), kx.size() * 4));
        RCC(up(&c->p_by, by.data(), by.size() * 4)); RCC(up(&c->p_ky, ky.data(), ky.size() * 4));
        RCC(up(&c->p_lut, lut.data(), lut.size() * 4));
        RCC(calloc_dev(c->pptrs, (void**)&c->p_tmp, (size_t)B * H * c->p_rw * 3));
        memcpy(c->p_mean, mean, 12);
        memcpy(c->p_std, std_, 12);
        c->pB = B; c->pH = H; c->pW = W;
    }
    const size_t n1 = (size_t)B * H * c->p_rw, n2 = (size_t)B * S * S;
    LTT_CUDA_OK(launch_k(clip_resize_h_kernel, dim3((unsigned)std::min<size_t>((n1 + 255) / 256, (size_t)c->sms * 16)), dim3(256), 0, st,
                         images, B, H, W, c->p_rw, (const int*)c->p_bx, (const int*)c->p_kx, c->p_ksx, c->p_tmp));
    LTT_CUDA_OK(launch_k(clip_resize_v_kernel, dim3((unsigned)std::min<size_t>((n2 + 255) / 256, (size_t)c->sms * 16)), dim3(256), 0, st,
                         (const uint8_t*)c->p_tmp, B, H, c->p_rw, S, c->p_top, c->p_left, (const int*)c->p_by, (const int*)c->p_ky,
                         c->p_ksy, (const float*)c->p_lut, pixel_values));
    c->launches += 2;
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t ltt_clip_vision_launch_count(const ltt_clip_vision* c) { return c ? c->launches : 0; }

// ---------------------------------------------------------------------------------------------------- reward head
int ltt_reward_head(const float* txt, const float* pred, const float* gt, int B, int D, const float* const* aes_w,
                    const float* const* aes_b, const int* aes_dims, const float* miou, const float* laysim, float* reward,
                    float* clip_reward, float* aes_reward, void* stream) {
    if (!txt || !pred || !gt || !aes_w || !aes_b || !aes_dims || !reward || B < 1 || D < 1 || D > 1024) {
        set_error("ltt_reward_head: bad arguments (B=%d, D=%d; D <= 1024)", B, D);
        return -1;
    }
    AesW aw;
    for (int i = 0; i < 6; ++i) {
        aw.dim[i] = aes_dims[i];
        if (aes_dims[i] < 1 || aes_dims[i] > 1024 || (i < 5 && aes_dims[i] % 4)) {
            set_error("ltt_reward_head: aesthetic layer width %d outside [1, 1024] (input widths: multiples of 4)", aes_dims[i]);
            return -1;
        }
    }
    if (aw.dim[0] != D || aw.dim[5] != 1) {
        set_error("ltt_reward_head: aesthetic MLP must map D=%d -> 1 (got %d -> %d)", D, aw.dim[0], aw.dim[5]);
        return -1;
    }
    for (int i = 0; i < 5; ++i) {
        aw.w[i] = aes_w[i];
        aw.b[i] = aes_b[i];
        if (!aw.w[i] || !aw.b[i]) {
            set_error("ltt_reward_head: null aesthetic layer %d", i);
            return -1;
        }
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(RH_CLUSTER, B, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = RH_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, reward_head_kernel, txt, pred, gt, D, aw, miou, laysim, reward, clip_reward, aes_reward));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
