// CLIP text tower on the sm_100a kernels: the conditioning prep of the sampler.
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

// pooled[b, :] = hidden[b, e_b, :] with e_b = argmax(ids[b]) (first maximum; eos_token_id == 2) or the first position whose
// id == eos_token_id; text_embeds[b, p] = sum_k pooled[b, k] * proj[p, k] in fp32 (CLIPModel.text_projection, no bias).
// One CTA per sequence.
__global__ void __launch_bounds__(256) clip_pool_kernel(const int32_t* __restrict__ ids, const float* __restrict__ hidden, int L,
                                                        int W, int eos_token_id, const float* __restrict__ proj, int P,
                                                        float* __restrict__ pooled, float* __restrict__ embeds) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float prow[];
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
        if (pooled) pooled[(size_t)b * W + i] = v;
    }
    __syncthreads();
    if (!embeds) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = warp; p < P; p += blockDim.x >> 5) {
        const float* w = proj + (size_t)p * W;
        float acc = 0.f;
        for (int k = lane; k < W; k += 32) acc = fmaf(prow[k], w[k], acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) embeds[(size_t)b * P + p] = acc;
    }
}

struct CParam { float* dev = nullptr; std::vector<int64_t> shape; size_t numel = 0; };
struct CLayer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    __half *wqkv, *wo, *w1, *w2;
    const float *bqkv, *bo, *b1, *b2;
};

}  // namespace ltt

using namespace ltt;

struct ltt_clip {
    ltt_clip_config cfg;
    int device = 0, sms = 148;
    std::map<std::string, CParam> params;
    bool finalized = false;
    std::vector<void*> wptrs, cptrs;
    std::vector<CLayer> layers;
    const float *tok = nullptr, *pos = nullptr, *fin_g = nullptr, *fin_b = nullptr, *proj = nullptr;
    // workspace for (B, L)
    int B = 0, L = 0, pitch_v = 0;
    float *x = nullptr, *hid = nullptr;
    __half *h16 = nullptr, *q = nullptr, *k = nullptr, *vt = nullptr, *a16 = nullptr, *f16 = nullptr;
    int64_t launches = 0;
};

namespace ltt {

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
static const CParam* cfind(ltt_clip* c, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = c->params.find(key);
    if (it == c->params.end()) {
        set_error("clip: missing parameter '%s' (load_state_dict incomplete)", key.c_str());
        return nullptr;
    }
    if (it->second.shape != std::vector<int64_t>(shape)) {
        set_error("clip: parameter '%s' has the wrong shape for this configuration", key.c_str());
        return nullptr;
    }
    return &it->second;
}
#define CGET(var, key, ...)                              \
    const CParam* var = cfind(c, (key), {__VA_ARGS__});  \
    if (!var) return -6;

static int clip_build(ltt_clip* c) {
    const ltt_clip_config& g = c->cfg;
    const int64_t W = g.hidden, F = g.ffn;
    crelease(c->wptrs);
    c->layers.clear();
    const std::string tm = "text_model.";
    {
        CGET(t, tm + "embeddings.token_embedding.weight", g.vocab, W)
        CGET(p, tm + "embeddings.position_embedding.weight", g.max_pos, W)
        CGET(fg, tm + "final_layer_norm.weight", W)
        CGET(fb, tm + "final_layer_norm.bias", W)
        c->tok = t->dev; c->pos = p->dev; c->fin_g = fg->dev; c->fin_b = fb->dev;
    }
    c->proj = nullptr;
    if (g.proj_dim > 0) {
        CGET(pw, "text_projection.weight", g.proj_dim, W)
        c->proj = pw->dev;
    }
    for (int i = 0; i < g.layers; ++i) {
        const std::string p = tm + "encoder.layers." + std::to_string(i) + ".";
        CLayer l{};
        CGET(g1, p + "layer_norm1.weight", W) CGET(b1, p + "layer_norm1.bias", W)
        CGET(g2, p + "layer_norm2.weight", W) CGET(b2, p + "layer_norm2.bias", W)
        l.ln1_g = g1->dev; l.ln1_b = b1->dev; l.ln2_g = g2->dev; l.ln2_b = b2->dev;
        void *wq, *bq;
        RCC(calloc_dev(c->wptrs, &wq, (size_t)3 * W * W * 2));
        RCC(calloc_dev(c->wptrs, &bq, (size_t)3 * W * 4));
        const char* names[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; ++j) {       // q / k / v stacked into one [3W, W] matrix (the softmax scale stays in the attention kernel)
            CGET(w, p + "self_attn." + names[j] + ".weight", W, W)
            CGET(b, p + "self_attn." + names[j] + ".bias", W)
            RCC(pack_rows_launch(w->dev, (int)W, (int)W, (__half*)wq, j * (int)W, 0, 0));
            LTT_CUDA_OK(cudaMemcpy((float*)bq + j * W, b->dev, W * 4, cudaMemcpyDeviceToDevice));
        }
        l.wqkv = (__half*)wq; l.bqkv = (const float*)bq;
        auto lin = [&](const std::string& name, int64_t N, int64_t K, __half** wout, const float** bout) -> int {
            CGET(w, p + name + ".weight", N, K)
            CGET(b, p + name + ".bias", N)
            void* q;
            RCC(calloc_dev(c->wptrs, &q, (size_t)N * K * 2));
            RCC(pack_rows_launch(w->dev, (int)N, (int)K, (__half*)q, 0, 0, 0));
            *wout = (__half*)q; *bout = b->dev;
            return 0;
        };
        RCC(lin("self_attn.out_proj", W, W, &l.wo, &l.bo));
        RCC(lin("mlp.fc1", F, W, &l.w1, &l.b1));
        RCC(lin("mlp.fc2", W, F, &l.w2, &l.b2));
        c->layers.push_back(l);
    }
    LTT_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

static int clip_workspace(ltt_clip* c, int B, int L) {
    crelease(c->cptrs);
    const ltt_clip_config& g = c->cfg;
    const size_t M = (size_t)B * L, W = g.hidden;
    c->pitch_v = (L + 7) & ~7;
    auto A = [&](auto** p, size_t bytes) { return calloc_dev(c->cptrs, (void**)p, bytes); };
    RCC(A(&c->x, M * W * 4)); RCC(A(&c->hid, M * W * 4));
    RCC(A(&c->h16, M * W * 2)); RCC(A(&c->q, M * W * 2)); RCC(A(&c->k, M * W * 2)); RCC(A(&c->a16, M * W * 2));
    RCC(A(&c->vt, (size_t)B * W * c->pitch_v * 2));
    RCC(A(&c->f16, M * (size_t)g.ffn * 2));
    LTT_CUDA_OK(cudaMemset(c->vt, 0, (size_t)B * W * c->pitch_v * 2));      // pitch padding columns are never written
    c->B = B; c->L = L;
    return 0;
}

// rows = nb sequences x len tokens; the QKV projection passes the real geometry (its epilogue transposes V per sequence),
// the other GEMMs run on the flat [B * L] row list (nb = 1)
static int clip_gemm(ltt_clip* c, cudaStream_t st, int nb, int len, int N, const __half* a, int K, const __half* w,
                     const GemmEpilogue& epi) {
    GemmProblem p{};
    p.B = nb; p.H = 1; p.W = len; p.N = N; p.nsrc = 1;
    p.src[0] = GemmSrc{a, K, K, 1};
    p.w = w; p.Ktot = K; p.w_static = 1; p.epi = epi;
    c->launches++;
    return gemm_tc_launch(p, c->sms, st);
}

}  // namespace ltt

extern "C" {

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
    LTT_CUDA_OK(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    *out = c;
    return 0;
}

void ltt_clip_destroy(ltt_clip* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    crelease(c->wptrs);
    crelease(c->cptrs);
    for (auto& kv : c->params) cudaFree(kv.second.dev);
    delete c;
}

int ltt_clip_load_param(ltt_clip* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_host) {
    if (!c || !key || !data || (ndim > 0 && !shape)) {
        set_error("ltt_clip_load_param: null argument");
        return -1;
    }
    LTT_CUDA_OK(cudaSetDevice(c->device));
    CParam& p = c->params[key];
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
    c->finalized = false;
    return 0;
}

int ltt_clip_finalize(ltt_clip* c) {
    if (!c) return -1;
    LTT_CUDA_OK(cudaSetDevice(c->device));
    RCC(clip_build(c));
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
    if (B != c->B || L != c->L) {
        LTT_CUDA_OK(cudaDeviceSynchronize());
        RCC(clip_workspace(c, B, L));
    }
    const int M = B * L, W = g.hidden, H = g.heads;
    {
        const int blocks = (int)std::min<size_t>(((size_t)M * (W >> 2) + 255) / 256, (size_t)c->sms * 8);
        LTT_CUDA_OK(launch_k(clip_embed_kernel, dim3(blocks), dim3(256), 0, st, ids, c->tok, c->pos, M, L, W, g.vocab, c->x));
        c->launches++;
    }
    for (const CLayer& l : c->layers) {
        RCC(layernorm_launch(c->x, DT_F32, M, W, l.ln1_g, l.ln1_b, g.eps, c->h16, nullptr, st));
        c->launches++;
        {
            GemmEpilogue e;
            e.bias = l.bqkv;
            e.out_mode = OUT_QKV;
            e.q = c->q; e.k = c->k; e.vt = c->vt;
            e.C = W; e.dhead = 64; e.dpad = 64; e.rows_q = L; e.rows_k = L; e.pitch_v = c->pitch_v; e.tokens = L; e.qkv_base = 0;
            RCC(clip_gemm(c, st, B, L, 3 * W, c->h16, W, l.wqkv, e));
        }
        {
            AttnProblem p{};
            p.B = B; p.heads = H; p.dhead = 64; p.dpad = 64; p.nq = L; p.nk = L;
            p.q = c->q; p.rows_q = L; p.k = c->k; p.rows_k = L; p.vt = c->vt; p.pitch_v = c->pitch_v;
            p.out = c->a16; p.ldo = W; p.scale = 0.125f; p.causal = 1;
            RCC(attn_tc_launch(p, st));
            c->launches++;
        }
        {
            GemmEpilogue e;
            e.bias = l.bo; e.res = c->x; e.res_dtype = DT_F32; e.ldr = W; e.out = c->x; e.out_dtype = DT_F32; e.ldo = W;
            RCC(clip_gemm(c, st, 1, M, W, c->a16, W, l.wo, e));
        }
        RCC(layernorm_launch(c->x, DT_F32, M, W, l.ln2_g, l.ln2_b, g.eps, c->h16, nullptr, st));
        c->launches++;
        {
            GemmEpilogue e;
            e.bias = l.b1; e.act = ACT_QUICKGELU; e.out = c->f16; e.out_dtype = DT_F16; e.ldo = g.ffn;
            RCC(clip_gemm(c, st, 1, M, g.ffn, c->h16, W, l.w1, e));
        }
        {
            GemmEpilogue e;
            e.bias = l.b2; e.res = c->x; e.res_dtype = DT_F32; e.ldr = W; e.out = c->x; e.out_dtype = DT_F32; e.ldo = W;
            RCC(clip_gemm(c, st, 1, M, W, c->f16, g.ffn, l.w2, e));
        }
    }
    float* hid = last_hidden ? last_hidden : c->hid;
    RCC(layernorm_launch(c->x, DT_F32, M, W, c->fin_g, c->fin_b, g.eps, nullptr, hid, st));
    c->launches++;
    if (pooled || text_embeds) {
        LTT_CUDA_OK(launch_k(clip_pool_kernel, dim3(B), dim3(256), (size_t)W * 4, st, ids, (const float*)hid, L, W, g.eos_token_id,
                             text_embeds ? c->proj : (const float*)nullptr, g.proj_dim, pooled, text_embeds));
        c->launches++;
    }
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int64_t ltt_clip_launch_count(const ltt_clip* c) { return c ? c->launches : 0; }

}  // extern "C"
