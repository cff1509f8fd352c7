// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[m, n] = sum_k A[m, k] * Wp[n, k]      fp16 x fp16 -> fp32 in TMEM
//
// One CTA computes 128 x BN output tiles (persistent, accumulator double buffered in TMEM).  Warp roles: warp 0 = TMA
// producer, warp 1 = TMEM owner + tcgen05.mma issuer (both loops run warp-converged, one elected lane issues), warps
// 2.. = EPI_WGS epilogue warpgroups (TMEM -> registers -> fused epilogue -> global).  A-tiles are fetched by 4-D TMA
// boxes (64 channels x tw x th x tb pixels = 128 rows) straight out of the NHWC activation; the nine taps of a 3x3
// convolution are nine shifted boxes accumulated into the same TMEM tile, padding comes from TMA's out-of-bounds zero
// fill, so no im2col buffer ever exists.  Channel-concatenated inputs and the ResBlock's 1x1 skip convolution are
// extra A sources appended along K.  splits > 1 = split-K over a thread-block cluster: the K range is divided over the
// cluster ranks, partial accumulators meet in distributed shared memory and are summed in fixed rank order.
// PAIR = CTA-pair mode (cta_group::2): a 2-CTA cluster computes 256 x BN tiles, each CTA stages half of the B tile.
//
// Replaces (reference, /root/reference/GLIGEN/ldm/modules): every nn.Linear / nn.Conv2d on the UNet path --
// attention.py:108-112,153-157 (q/k/v/out projections), :38-65 (GEGLU feed-forward), :425-433 (proj_in/out),
// diffusionmodules/openaimodel.py:155-194 (ResBlock convs, emb_layers, skip), :57-114 (Up/Downsample convs).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "ltt_kernels.h"
#include "ltt_ptx.cuh"

namespace ltt {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;   // 16 KB
#ifndef LTT_EPI_WGS
#define LTT_EPI_WGS 3
#endif
constexpr int EPI_WGS = LTT_EPI_WGS;                          // epilogue warpgroups (the epilogue is latency bound: more warps)
// Experiment build (-DLTT_LITE=1 -DLTT_EPI_WGS=1): CTAs small enough for TWO per SM (192 threads, <= 96 KB of shared memory,
// 256 TMEM columns), so that consecutive GEMM launches overlap under PDL; tile widths 64 / 128 only.
#ifndef LTT_LITE
#define LTT_LITE 0
#endif
constexpr int GEMM_MINB = LTT_LITE ? 2 : 1;
constexpr int ST64 = LTT_LITE ? 3 : 6, ST128 = LTT_LITE ? 3 : 6, ST160 = LTT_LITE ? 2 : 5, ST256 = LTT_LITE ? 2 : 4;
constexpr int EPI_THREADS = 128 * EPI_WGS;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;      // TMA warp, MMA warp, epilogue warpgroups

struct GemmDeviceArgs {
    CUtensorMap amap[3];
    CUtensorMap bmap;
    int nsrc;
    int taps[3];
    int kchunks[3];
    int iters_total;
    int B, H, W, N;
    int tw, th, tiles_x, tiles_y;
    int kind;                     // epilogue_kind(epi)
    int mtiles, ntiles, splits;   // splits > 1: one cluster of `splits` CTAs per tile, K range split by cluster rank
    int w_static;                 // weights may be fetched before the predecessor grid completes
    int pair;                     // 1: CTA-pair mode (cta_group::2): cluster of 2 CTAs computes a 256 x BN tile, each CTA
                                  // loads its own 128 A rows and half of the B tile, the leader issues 256-row MMAs
    GemmEpilogue epi;
};

template <int BN, int STAGES>
struct GemmSmem {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int PAIR_STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES / 2;     // per CTA of a pair
    static constexpr int BIAS_OFFSET = STAGES * STAGE_BYTES;           // BN floats
    static constexpr int LNS_OFFSET = BIAS_OFFSET + BN * 4;            // BN floats: ln_s of this tile (LayerNorm fold)
    static constexpr int BAR_OFFSET = LNS_OFFSET + BN * 4;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;              // + barriers/tmem slot + alignment slack
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// GELU (erf form) through the Abramowitz-Stegun 7.1.26 erfc approximation (|abs err| < 5e-7 on x * Phi(x)): after the
// fp16 rounding the reference applies to the GELU output it differs from the exact value on fewer fp16 inputs (254 of
// 63488) than the erff() formulation does (333), at about half the instructions.
__device__ __forceinline__ float gelu_erf_f(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    p *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float tail = 0.5f * p * e;                 // Phi(-|x|)
    return x * (x >= 0.f ? 1.0f - tail : tail);
}
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

struct RowInfo {
    int m;       // global row
    int b;       // batch element
    bool valid;
};

// running (sum, sum of squares) of the fp16 values a thread stored for its row (producer side of the LayerNorm fold)
struct RowStats {
    float s = 0.f, q = 0.f;
    __device__ __forceinline__ void add(const __half2* h) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            s += f.x + f.y;
            q = fmaf(f.x, f.x, fmaf(f.y, f.y, q));
        }
    }
};

// Residual / time-embedding operands of one row for 16 consecutive output columns, fetched BEFORE the accumulator
// is read so their latency hides behind the TMEM load (and behind the previous chunk's arithmetic).
struct RowOperands {
    uint4 res[4];      // fp16: res[0..1] hold 16 halves; fp32: res[0..3] hold 16 floats
    uint4 rv[2];       // 16 halves of the per-batch row vector
};

__device__ __forceinline__ void fetch_operands(const GemmEpilogue& e, const RowInfo& ri, int n, int Nout, RowOperands& o) {
    if (!ri.valid || n >= Nout || e.out_mode == OUT_QKV) return;
    // Nout is a multiple of 8; a 16-column chunk may end past Nout only in whole groups of 8
    if (e.res) {
        if (e.res_dtype == DT_F16) {
            const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(e.res) + (size_t)ri.m * e.ldr + n);
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (n + 8 * i < Nout) o.res[i] = p[i];
        } else {
            const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.res) + (size_t)ri.m * e.ldr + n);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (n + 4 * i < Nout) o.res[i] = p[i];
        }
    }
    if (e.rowvec) {
        const uint4* p = reinterpret_cast<const uint4*>(e.rowvec + (size_t)ri.b * e.ld_rowvec + n);
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (n + 8 * i < Nout) o.rv[i] = p[i];
    }
}

// Fused epilogue on 8 consecutive output columns [n, n+8) of row `ri`: v = raw fp32 accumulators (GEGLU: value half,
// g = gate half), bv / bg = the matching bias values (already in registers), j = index of the group inside the
// 16-column chunk whose operands are in `o`.  Rounding points follow the fp16-autocast reference: each Linear/Conv
// output is rounded to fp16 before the next elementwise op.
template <bool LNF>
__device__ __forceinline__ void epilogue_group8(const GemmEpilogue& e, const RowInfo& ri, int n, int Nout, float (&v)[8],
                                                const float (&g)[8], const float (&bv)[8], const float (&bg)[8],
                                                const RowOperands& o, int j, RowStats& rs) {
    if (!ri.valid || n >= Nout) return;
    if (e.act == ACT_GEGLU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float val = r16(v[i] + bv[i]);
            const float gat = r16(g[i] + bg[i]);
            v[i] = r16(val * r16(gelu_erf_f(gat)));
        }
    } else {
        const __half2* rvh = reinterpret_cast<const __half2*>(&o.rv[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = v[i] + bv[i];
            if (e.out_dtype == DT_F16 || e.res || e.rowvec || e.act) y = r16(y);
            if (e.rowvec) {
                const float2 f = __half22float2(rvh[i >> 1]);
                y = r16(y + ((i & 1) ? f.y : f.x));
            }
            if (e.act == ACT_SILU) y = r16(silu_f(y));
            v[i] = y;
        }
    }
    if (e.has_gate) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = r16(e.gate * v[i]);
    }
    if (e.out_mode == OUT_QKV) {
        const int which = n / e.C + e.qkv_base, c = n % e.C;
        const int t = ri.m - ri.b * e.tokens;
        if (which == 2) {
            __half* dst = e.vt + ((size_t)ri.b * e.C + c) * e.pitch_v + t;
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[(size_t)i * e.pitch_v] = __float2half_rn(v[i]);
        } else {
            const int head = c / e.dhead, jj = c % e.dhead;
            __half* base = which == 0 ? e.q + ((size_t)ri.b * e.rows_q + t) * (size_t)(e.dpad * (e.C / e.dhead))
                                      : e.k + ((size_t)ri.b * e.rows_k + t) * (size_t)(e.dpad * (e.C / e.dhead));
            __half2 h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            *reinterpret_cast<uint4*>(base + head * e.dpad + jj) = *reinterpret_cast<uint4*>(h);
        }
        return;
    }
    if (e.res) {
        if (e.res_dtype == DT_F16) {
            const __half2* rh = reinterpret_cast<const __half2*>(&o.res[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(rh[i]);
                v[2 * i] += f.x;
                v[2 * i + 1] += f.y;
            }
        } else {
            const float4 a = *reinterpret_cast<const float4*>(&o.res[2 * j]);
            const float4 b4 = *reinterpret_cast<const float4*>(&o.res[2 * j + 1]);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            v[4] += b4.x; v[5] += b4.y; v[6] += b4.z; v[7] += b4.w;
        }
    }
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    if constexpr (LNF) {
        if (e.stats_out) rs.add(h);
    }
    if (e.out_dtype == DT_F16) {
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + (size_t)ri.m * e.ldo + n) = *reinterpret_cast<uint4*>(h);
    } else {
        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)ri.m * e.ldo + n);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
        if constexpr (LNF) {
            if (e.out16) *reinterpret_cast<uint4*>(e.out16 + (size_t)ri.m * e.ldo + n) = *reinterpret_cast<uint4*>(h);
        }
    }
}

// ---- specialised epilogues: the model's GEMMs fall into a handful of epilogue shapes; each gets a branch-free inner
// loop (the generic routine above costs ~12 issue slots per element, which made small-K GEMMs epilogue bound).
// All produce bit-identical results to epilogue_group8 for their flag combination.
enum : int { EK_GENERIC = 0, EK_F16 = 1, EK_GATE16 = 2, EK_GEGLU = 3, EK_QKV = 4, EK_RES32 = 5, EK_QGELU = 6 };

__host__ __device__ inline int epilogue_kind(const GemmEpilogue& e) {
    if (e.out_mode == OUT_QKV) return (!e.act && !e.has_gate && !e.rowvec) ? EK_QKV : EK_GENERIC;
    if (e.act == ACT_GEGLU) return (!e.res && !e.has_gate && !e.rowvec && e.out_dtype == DT_F16) ? EK_GEGLU : EK_GENERIC;
    if (e.act == ACT_QUICKGELU) return EK_QGELU;      // the launcher rejects any other operand with it
    if (e.act != ACT_NONE || e.rowvec) return EK_GENERIC;
    if (e.res && e.res_dtype == DT_F32) return e.has_gate ? EK_GENERIC : EK_RES32;
    if (e.out_dtype != DT_F16) return EK_GENERIC;
    if (e.has_gate) return e.res ? EK_GATE16 : EK_GENERIC;
    return EK_F16;      // bias -> fp16, optional fp16 residual
}

template <int K>
struct KindTag {
    static constexpr int value = K;
};

template <int KIND, bool LNF>
__device__ __forceinline__ void epi_group8(const GemmEpilogue& e, const RowInfo& ri, int n, int Nout, float (&v)[8],
                                           const float (&g)[8], const float* __restrict__ bs, const float* __restrict__ bsg,
                                           const RowOperands& o, int j, RowStats& rs) {
    if constexpr (KIND == EK_GENERIC) {
        float bv[8], bg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bv[i] = bs[i];
            bg[i] = e.act == ACT_GEGLU ? bsg[i] : 0.f;
        }
        epilogue_group8<LNF>(e, ri, n, Nout, v, g, bv, bg, o, j, rs);
        return;
    } else {
        if (!ri.valid || n >= Nout) return;
        __half2 h[4];
        if constexpr (KIND == EK_F16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i] + bs[2 * i], v[2 * i + 1] + bs[2 * i + 1]);
            if (e.res) {
                const __half2* rh = reinterpret_cast<const __half2*>(&o.res[j]);
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __hadd2(h[i], rh[i]);     // one rounding of the exact sum
            }
        } else if constexpr (KIND == EK_QGELU) {        // CLIP MLP: x * sigmoid(1.702 x) on the fp32 value (fp32 reference), fp16 out
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float y0 = v[2 * i] + bs[2 * i], y1 = v[2 * i + 1] + bs[2 * i + 1];
                h[i] = __floats2half2_rn(y0 / (1.0f + __expf(-1.702f * y0)), y1 / (1.0f + __expf(-1.702f * y1)));
            }
        } else if constexpr (KIND == EK_GATE16) {
            const __half2* rh = reinterpret_cast<const __half2*>(&o.res[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 y = __half22float2(__floats2half2_rn(v[2 * i] + bs[2 * i], v[2 * i + 1] + bs[2 * i + 1]));
                h[i] = __hadd2(__floats2half2_rn(e.gate * y.x, e.gate * y.y), rh[i]);
            }
        } else if constexpr (KIND == EK_GEGLU) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __half2 val = __floats2half2_rn(v[2 * i] + bs[2 * i], v[2 * i + 1] + bs[2 * i + 1]);
                const float2 gt = __half22float2(__floats2half2_rn(g[2 * i] + bsg[2 * i], g[2 * i + 1] + bsg[2 * i + 1]));
                h[i] = __hmul2(val, __floats2half2_rn(gelu_erf_f(gt.x), gelu_erf_f(gt.y)));
            }
        } else if constexpr (KIND == EK_QKV) {
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i] + bs[2 * i], v[2 * i + 1] + bs[2 * i + 1]);   // bias: 0 unless LN fold / conv bias
            const int which = n / e.C + e.qkv_base, c = n % e.C;
            const int t = ri.m - ri.b * e.tokens;
            if (which == 2) {
                __half* dst = e.vt + ((size_t)ri.b * e.C + c) * e.pitch_v + t;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dst[(size_t)(2 * i) * e.pitch_v] = __low2half(h[i]);
                    dst[(size_t)(2 * i + 1) * e.pitch_v] = __high2half(h[i]);
                }
            } else {
                const int head = c / e.dhead, jj = c % e.dhead;
                __half* base = which == 0 ? e.q + ((size_t)ri.b * e.rows_q + t) * (size_t)(e.dpad * (e.C / e.dhead))
                                          : e.k + ((size_t)ri.b * e.rows_k + t) * (size_t)(e.dpad * (e.C / e.dhead));
                *reinterpret_cast<uint4*>(base + head * e.dpad + jj) = *reinterpret_cast<uint4*>(h);
            }
            return;
        } else if constexpr (KIND == EK_RES32) {
            const float4 a = *reinterpret_cast<const float4*>(&o.res[2 * j]);
            const float4 b4 = *reinterpret_cast<const float4*>(&o.res[2 * j + 1]);
            const float rr[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
            float y[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(__floats2half2_rn(v[2 * i] + bs[2 * i], v[2 * i + 1] + bs[2 * i + 1]));
                y[2 * i] = f.x + rr[2 * i];
                y[2 * i + 1] = f.y + rr[2 * i + 1];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            if (e.out_dtype == DT_F32) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)ri.m * e.ldo + n);
                op[0] = make_float4(y[0], y[1], y[2], y[3]);
                op[1] = make_float4(y[4], y[5], y[6], y[7]);
                if constexpr (LNF) {
                    if (e.stats_out) rs.add(h);
                    if (e.out16) *reinterpret_cast<uint4*>(e.out16 + (size_t)ri.m * e.ldo + n) = *reinterpret_cast<uint4*>(h);
                }
                return;
            }
        }
        if constexpr (LNF && (KIND == EK_F16 || KIND == EK_GATE16 || KIND == EK_RES32)) {
            if (e.stats_out) rs.add(h);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out) + (size_t)ri.m * e.ldo + n) = *reinterpret_cast<uint4*>(h);
    }
}

// Persistent, warp-specialised kernel.  Work unit = (output tile 128 x BN, K split).  Each CTA walks units
// blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA producer and the MMA issuer run ahead of the epilogue warps by up
// to one unit because the fp32 accumulator is double buffered in TMEM (2 x BN columns): the epilogue of unit u
// overlaps the main loop of unit u + 1.
// LNF: the LayerNorm-fold features (row statistics out, fp16 copy of an fp32 output, folded-LayerNorm consumer); the plain
// variant carries none of that code.
template <int BN, int STAGES, bool PAIR, bool LNF>
__global__ void __launch_bounds__(GEMM_THREADS, GEMM_MINB) gemm_tc_kernel(const __grid_constant__ GemmDeviceArgs args) {
    using SM = GemmSmem<BN, STAGES>;
    constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* bias_s = reinterpret_cast<float*>(smem + SM::BIAS_OFFSET);
    float* lns_s = reinterpret_cast<float*>(smem + SM::LNS_OFFSET);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;     // 2
    uint64_t* acc_empty = acc_full + 2;          // 2 (count EPI_THREADS)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    // warp index through a shuffle so ptxas treats it (and every branch on it) as warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // CTA-pair mode: work unit = (pair of M tiles, N tile); CTA rank r of the cluster owns M tile 2 * mpair + r.
    constexpr bool pair = PAIR;     // a kernel containing cta_group::2 instructions can only be launched as a cluster
    const uint32_t crank = pair ? cluster_ctarank() : 0u;
    const int cta_step = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;     // units advance by clusters in pair mode
    const int cta_first = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    // Split-K mode (splits > 1): the grid is exactly one cluster per tile, every CTA runs ONE unit (tile = cluster id,
    // K range = cluster rank) and the partial accumulators are reduced through distributed shared memory.
    const int total_units = pair ? ((args.mtiles + 1) >> 1) * args.ntiles : args.mtiles * args.ntiles * args.splits;
    const int per = (args.iters_total + args.splits - 1) / args.splits;   // host guarantees every z gets >= 1 iteration
    const int tiles_img = args.tiles_x * args.tiles_y;
    const int tb = BM / (args.tw * args.th);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < args.nsrc; ++s) tma_prefetch_desc(&args.amap[s]);
        tma_prefetch_desc(&args.bmap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], pair ? 2 * EPI_THREADS : EPI_THREADS);   // pair: both CTAs' epilogues report to the leader
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if constexpr (pair) tmem_alloc2<TMEM_COLS>(tmem_slot);
        else tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();      // the peer's barriers are initialised before anything is signalled remotely
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    // everything above overlapped the predecessor's tail (programmatic dependent launch); from here on global memory
    // written by it is read
    pdl_launch_dependents();
    // The producer and MMA-issuer warps run their loops CONVERGED with warp-uniform operands; a single elected lane
    // performs each TMA / tcgen05 operation inside the "_elect" wrappers (see ltt_ptx.cuh).
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_addr = smem_u32(full_bar), empty_addr = smem_u32(empty_bar), accfull_addr = smem_u32(acc_full);
    // Weights are not produced by the predecessor (w_static): the producer starts streaming the B tiles of its first
    // unit's first STAGES iterations BEFORE waiting on the predecessor grid, so the HBM latency of the weight stream
    // hides behind the predecessor's tail.  Only the activation (A) loads wait.
    int npre = 0;
    if (warp == 0 && args.w_static && !pair && (int)blockIdx.x < total_units) {
        const int unit = blockIdx.x;
        const int z = unit % args.splits, tile = unit / args.splits;
        const int n0 = (tile % args.ntiles) * BN;
        const int it0 = z * per, it1 = min(args.iters_total, it0 + per);
        npre = min(STAGES, it1 - it0);
        for (int i = 0; i < npre; ++i) {
            mbar_expect_tx_elect(full_addr + 8u * i, SM::STAGE_BYTES);
            tma_load_2d_elect(smem_base + i * SM::STAGE_BYTES + A_TILE_BYTES, &args.bmap, full_addr + 8u * i, (it0 + i) * BK, n0);
        }
    }
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full_leader = pair ? dsmem_map(full_addr, 0) : 0u;   // leader's full_bar[0] (cluster window)
        for (int unit = cta_first; unit < total_units; unit += cta_step) {
            const int z = pair ? 0 : unit % args.splits, tile = pair ? unit : unit / args.splits;
            const int n0 = (tile % args.ntiles) * BN;
            const int mt = pair ? (tile / args.ntiles) * 2 + (int)crank : tile / args.ntiles;
            const int tile_b = mt / tiles_img, trem = mt % tiles_img;
            const int b0 = tile_b * tb, y0 = (trem / args.tiles_x) * args.th, x0 = (trem % args.tiles_x) * args.tw;
            const int it0 = z * per, it1 = min(args.iters_total, it0 + per);
            // decode (source, tap, chunk) of the first iteration
            int s = 0, tap = 0, chunk = 0;
            {
                int rem = it0;
                while (s < args.nsrc - 1 && rem >= args.taps[s] * args.kchunks[s]) {
                    rem -= args.taps[s] * args.kchunks[s];
                    ++s;
                }
                tap = rem / args.kchunks[s];
                chunk = rem % args.kchunks[s];
            }
            for (int it = it0; it < it1; ++it) {
                const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
                int dx = 0, dy = 0;
                if (args.taps[s] == 9) {
                    dy = tap / 3 - 1;
                    dx = tap % 3 - 1;
                }
                const CUtensorMap* am = s == 0 ? &args.amap[0] : (s == 1 ? &args.amap[1] : &args.amap[2]);
                if constexpr (pair) {
                    // the stage is free in THIS CTA once the leader's multicast commit arrived here; both CTAs'
                    // bytes are credited to the leader's full barrier
                    mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
                    if (crank == 0) mbar_expect_tx_elect(full_addr + 8u * stage, 2 * SM::PAIR_STAGE_BYTES);
                    const uint32_t fb = full_leader + 8u * stage;
                    tma_load_4d_2sm_elect(sa, am, fb, chunk * BK, x0 + dx, y0 + dy, b0);
                    tma_load_2d_2sm_elect(sa + A_TILE_BYTES, &args.bmap, fb, it * BK, n0 + (int)crank * (BN / 2));
                } else {
                    const bool pre = npre > 0;     // B tile of this iteration is already in flight (see above)
                    if (!pre) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx_elect(full_addr + 8u * stage, SM::STAGE_BYTES);
                    } else {
                        --npre;
                    }
                    tma_load_4d_elect(sa, am, full_addr + 8u * stage, chunk * BK, x0 + dx, y0 + dy, b0);
                    if (!pre) tma_load_2d_elect(sa + A_TILE_BYTES, &args.bmap, full_addr + 8u * stage, it * BK, n0);
                }
                if (++chunk == args.kchunks[s]) {
                    chunk = 0;
                    if (++tap == args.taps[s]) {
                        tap = 0;
                        ++s;
                    }
                }
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        // pair mode: the leader CTA issues 256 x BN x 16 MMAs for both CTAs; its commits are multicast so the producer
        // (empty) and epilogue (acc_full) barriers of BOTH CTAs are released.
        if (!pair || crank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BN, pair ? 256 : 128);
            int stage = 0;
            uint32_t phase = 0;
            int u = 0;
            for (int unit = cta_first; unit < total_units; unit += cta_step, ++u) {
                const int z = pair ? 0 : unit % args.splits;
                const int it0 = z * per, it1 = min(args.iters_total, it0 + per);
                const int niter = it1 - it0;
                const int buf = u & 1;
                // epilogue(s) have drained this accumulator
                if constexpr (pair) mbar_wait_cluster(&acc_empty[buf], ((u >> 1) & 1) ^ 1);
                else mbar_wait(&acc_empty[buf], ((u >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * BN;
                for (int i = 0; i < niter; ++i) {
                    if constexpr (pair) mbar_wait_cluster(&full_bar[stage], phase);
                    else mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * SM::STAGE_BYTES;
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if constexpr (pair) umma_f16_2sm_elect(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                        else umma_f16_elect(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                    }
                    if constexpr (pair) {
                        umma_commit_2sm_elect(empty_addr + 8u * stage, 3);
                        if (i == niter - 1) umma_commit_2sm_elect(accfull_addr + 8u * buf, 3);
                    } else {
                        umma_commit_elect(empty_addr + 8u * stage);
                        if (i == niter - 1) umma_commit_elect(accfull_addr + 8u * buf);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue warps (2 .. 2 + 4*EPI_WGS): each
        // thread owns one tile row (TMEM lane) and its warpgroup's share of the 32-column chunks
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;            // tile row
        const int wg = (warp - 2) >> 2;         // 0 .. EPI_WGS-1
        const int et = threadIdx.x - 64;        // 0 .. EPI_THREADS-1
        const GemmEpilogue& e = args.epi;
        const bool geglu = e.act == ACT_GEGLU;
        const int Nout = geglu ? args.N / 2 : args.N;
        int u = 0;
        const uint32_t acc_empty_leader = pair ? dsmem_map(smem_u32(acc_empty), 0) : 0u;
        for (int unit = cta_first; unit < total_units; unit += cta_step, ++u) {
            const int z = pair ? 0 : unit % args.splits, tile = pair ? unit : unit / args.splits;
            const int n0 = (tile % args.ntiles) * BN;
            const int mt = pair ? (tile / args.ntiles) * 2 + (int)crank : tile / args.ntiles;
            const int buf = u & 1;
            RowInfo ri;
            {
                const int tile_b = mt / tiles_img, trem = mt % tiles_img;
                const int b0 = tile_b * tb, y0 = (trem / args.tiles_x) * args.th, x0 = (trem % args.tiles_x) * args.tw;
                const int pix = args.tw * args.th;
                const int ib = r / pix, rr = r % pix;
                const int iy = rr / args.tw, ix = rr % args.tw;
                const int b = b0 + ib, y = y0 + iy, x = x0 + ix;
                ri.valid = (b < args.B) && (y < args.H) && (x < args.W);
                ri.b = b;
                ri.m = (b * args.H + y) * args.W + x;
            }
            // output-column geometry of this tile.  GEGLU: the packed weight rows hold, per 128-row group, 64 value
            // rows followed by their 64 gate rows, so a BN tile carries BN/128 groups and BN/2 output columns.
            const int ncols = geglu ? BN / 2 : BN;
            const int nbase = geglu ? n0 / 2 : n0;
            // stage the bias of this tile in shared memory (value half, then gate half for GEGLU)
            named_bar_sync(2, EPI_THREADS);     // previous unit's readers are done with bias_s
            for (int c = et; c < BN; c += EPI_THREADS) {
                float bvv = 0.f;
                if (e.bias) {
                    if (geglu) {
                        const int half = c / ncols, cc = c % ncols;
                        if (nbase + cc < Nout) bvv = e.bias[half * Nout + nbase + cc];
                    } else if (n0 + c < Nout) {
                        bvv = e.bias[n0 + c];
                    }
                }
                bias_s[c] = bvv;
                if (LNF && e.ln_stats) {    // LayerNorm fold: ln_s is indexed like the bias
                    float sv = 0.f;
                    if (geglu) {
                        const int half = c / ncols, cc = c % ncols;
                        if (nbase + cc < Nout) sv = e.ln_s[half * Nout + nbase + cc];
                    } else if (n0 + c < Nout) {
                        sv = e.ln_s[n0 + c];
                    }
                    lns_s[c] = sv;
                }
            }
            RowOperands opA;
            if (args.splits == 1) fetch_operands(e, ri, nbase + 16 * wg, Nout, opA);
            // LayerNorm fold, consumer side: (mean, rstd) of this thread's row from the producer's partial sums -- fetched
            // while the main loop of this unit is still running
            float ln_mean = 0.f, ln_rstd = 1.f;
            if (LNF && e.ln_stats && ri.valid) {
                const float2* sp = e.ln_stats + (size_t)ri.m * e.ln_ld;
                if (e.ln_slots < 0) {
                    const float2 t = sp[0];
                    ln_mean = t.x;
                    ln_rstd = t.y;
                } else {
                    float sm = 0.f, sq = 0.f;
                    for (int i = 0; i < e.ln_slots; ++i) {      // fixed slot order: deterministic
                        const float2 t = sp[i];
                        sm += t.x;
                        sq += t.y;
                    }
                    const float inv = 1.0f / (float)e.ln_K;
                    ln_mean = sm * inv;
                    ln_rstd = rsqrtf(fmaxf(sq * inv - ln_mean * ln_mean, 0.f) + e.ln_eps);
                }
            }
            const float ln_nmr = -ln_mean * ln_rstd;       // y = rstd * acc + (-mean * rstd) * s_n + c_n
            RowStats rstat;
            named_bar_sync(2, EPI_THREADS);
            const uint32_t trow = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);

            if (pair) mbar_wait_cluster(&acc_full[buf], (u >> 1) & 1);
            else mbar_wait(&acc_full[buf], (u >> 1) & 1);
            tc_fence_after();

            if (args.splits > 1) {
                // ---- cluster split-K: dump the partial accumulator to this CTA's shared memory ([col][row] fp32, in the
                // idle pipeline stages: every TMA load has landed and every MMA has retired once acc_full fires), then
                // reduce a column slice over all ranks through DSMEM in fixed rank order (deterministic).
                float* part = reinterpret_cast<float*>(smem);
                fence_async_smem();
#pragma unroll 1
                for (int c = 32 * wg; c < BN; c += 32 * EPI_WGS) {
                    uint32_t v[32];
                    tmem_ld32(trow + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) part[(c + i) * BM + r] = __uint_as_float(v[i]);
                }
                tc_fence_before();
                cluster_sync_all();                                   // #1: all partial tiles are in shared memory
                const int S = args.splits;
                const int groups = ncols / 8;                         // 8-column output groups of this tile
                const int g0 = (groups * z) / S, g1 = (groups * (z + 1)) / S;
                const uint32_t part_addr = smem_u32(part);
                uint32_t peer[8];
#pragma unroll
                for (int zz = 0; zz < 8; ++zz) peer[zz] = zz < S ? dsmem_map(part_addr, zz) : 0u;
#pragma unroll 1
                for (int g = g0 + wg; g < g1; g += EPI_WGS) {
                    const int c = g * 8;
                    const int vcol = geglu ? (c / 64) * 128 + (c % 64) : c;
                    RowOperands o;
                    // operands of this 8-column group go to slot 0 of the chunk-shaped operand struct
                    if (ri.valid && nbase + c < Nout && e.out_mode != OUT_QKV) {
                        if (e.res) {
                            if (e.res_dtype == DT_F16) {
                                o.res[0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(e.res) + (size_t)ri.m * e.ldr + nbase + c);
                            } else {
                                const uint4* pp = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.res) + (size_t)ri.m * e.ldr + nbase + c);
                                o.res[0] = pp[0];
                                o.res[1] = pp[1];
                            }
                        }
                        if (e.rowvec) o.rv[0] = *reinterpret_cast<const uint4*>(e.rowvec + (size_t)ri.b * e.ld_rowvec + nbase + c);
                    }
                    float v8[8], g8[8], bv[8], bg[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v8[i] = 0.f;
                        g8[i] = 0.f;
                        bv[i] = bias_s[c + i];
                        bg[i] = geglu ? bias_s[ncols + c + i] : 0.f;
                    }
#pragma unroll
                    for (int zz = 0; zz < 8; ++zz) {
                        if (zz < S) {
                            float t8[8], u8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                t8[i] = dsmem_ld_f32(peer[zz] + (uint32_t)(((vcol + i) * BM + r) * 4));
                                u8[i] = geglu ? dsmem_ld_f32(peer[zz] + (uint32_t)(((vcol + 64 + i) * BM + r) * 4)) : 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                v8[i] += t8[i];
                                g8[i] += u8[i];
                            }
                        }
                    }
                    if (LNF && e.ln_stats) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v8[i] = fmaf(v8[i], ln_rstd, ln_nmr * lns_s[c + i]);
                            if (geglu) g8[i] = fmaf(g8[i], ln_rstd, ln_nmr * lns_s[ncols + c + i]);
                        }
                    }
                    epilogue_group8<LNF>(e, ri, nbase + c, Nout, v8, g8, bv, bg, o, 0, rstat);
                }
            } else {
                // 16-column chunks: accumulators -> fused epilogue (specialised per epilogue kind) -> global
                auto run_tile = [&](auto kind_tag) {
                    constexpr int KIND = decltype(kind_tag)::value;
                    constexpr bool kGeglu = KIND == EK_GEGLU;
                    constexpr bool kFetch = KIND == EK_GENERIC || KIND == EK_F16 || KIND == EK_GATE16 || KIND == EK_RES32;
                    bool fetched = true;        // operands of the first chunk were fetched before the accumulator wait
#pragma unroll 1
                    for (int c = 16 * wg; c < ncols; c += 16 * EPI_WGS) {
                        // column of the value / gate accumulator inside the tile for output column c + i
                        const bool gg = kGeglu || (KIND == EK_GENERIC && geglu);
                        const int vcol = gg ? (c / 64) * 128 + (c % 64) : c;
                        uint32_t v[16], g[16];
                        tmem_ld16(trow + vcol, v);
                        if (gg) tmem_ld16(trow + vcol + 64, g);
                        if (kFetch && !fetched) fetch_operands(e, ri, nbase + c, Nout, opA);   // overlaps the TMEM load
                        fetched = false;
                        tmem_ld_wait();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float v8[8], g8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                v8[i] = __uint_as_float(v[h * 8 + i]);
                                g8[i] = gg ? __uint_as_float(g[h * 8 + i]) : 0.f;
                            }
                            if (LNF && e.ln_stats) {
                                const float4* s4 = reinterpret_cast<const float4*>(lns_s + c + h * 8);
                                const float4 sa = s4[0], sb = s4[1];
                                const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) v8[i] = fmaf(v8[i], ln_rstd, ln_nmr * sv[i]);
                                if (gg) {
                                    const float4* g4 = reinterpret_cast<const float4*>(lns_s + ncols + c + h * 8);
                                    const float4 ga = g4[0], gb = g4[1];
                                    const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                                    for (int i = 0; i < 8; ++i) g8[i] = fmaf(g8[i], ln_rstd, ln_nmr * gv[i]);
                                }
                            }
                            epi_group8<KIND, LNF>(e, ri, nbase + c + h * 8, Nout, v8, g8, bias_s + c + h * 8, bias_s + ncols + c + h * 8, opA, h, rstat);
                        }
                    }
                };
                switch (args.kind) {
                    case EK_F16: run_tile(KindTag<EK_F16>{}); break;
                    case EK_GATE16: run_tile(KindTag<EK_GATE16>{}); break;
                    case EK_GEGLU: run_tile(KindTag<EK_GEGLU>{}); break;
                    case EK_QKV: run_tile(KindTag<EK_QKV>{}); break;
                    case EK_RES32: run_tile(KindTag<EK_RES32>{}); break;
                    case EK_QGELU: run_tile(KindTag<EK_QGELU>{}); break;
                    default: run_tile(KindTag<EK_GENERIC>{}); break;
                }
                tc_fence_before();
                if (pair) mbar_arrive_remote(acc_empty_leader + (uint32_t)buf * 8u);
                else mbar_arrive(&acc_empty[buf]);
            }
            if (LNF && e.stats_out && ri.valid) {
                const int slot = ((tile % args.ntiles) * args.splits + z) * EPI_WGS + wg;
                e.stats_out[(size_t)ri.m * e.stats_ld + slot] = make_float2(rstat.s, rstat.q);
            }
        }
    }
    __syncwarp();
    if (args.splits > 1) {
        if (warp < 2) cluster_sync_all();   // #1 (the epilogue warps arrived inside their role)
        cluster_sync_all();                 // #2: no CTA leaves while a peer may still read its partial tile
    }
    tc_fence_before();
    __syncthreads();
    if (pair) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair's MMAs / remote arrives are in flight
    if (warp == 1) {
        tc_fence_after();
        if constexpr (pair) tmem_dealloc2<TMEM_COLS>(tmem_base);
        else tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
static char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) on = getenv("LTT_NO_PDL") ? 0 : 1;
    return on != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
        return -3;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i - 1] * 2;
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                  rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
        return -4;
    }
    return 0;
}

static void pick_tile(int B, int H, int W, int* tw, int* th) {
    // 128 rows = tb images x th rows x tw pixels (powers of two); maximise the fraction of valid rows
    double best = -1.0;
    for (int w = 128; w >= 4; w >>= 1) {
        int h = 128 / w;
        while (h > 1 && h / 2 >= H) h >>= 1;
        const int b = 128 / (w * h);
        const double u = ((double)W / (((W + w - 1) / w) * w)) * ((double)H / (((H + h - 1) / h) * h)) *
                         ((double)B / (((B + b - 1) / b) * b));
        if (u > best + 1e-9) {
            best = u;
            *tw = w;
            *th = h;
        }
    }
}

// One kernel variant: picks the split-K cluster size and launches.
template <int BN, int STAGES>
struct Variant {
    using SM = GemmSmem<BN, STAGES>;
    static bool lnf(const GemmDeviceArgs& a) { return a.epi.stats_out || a.epi.ln_stats || a.epi.out16; }
    static int configure() {
        static bool configured = false;
        if (!configured) {
            LTT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
            LTT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
            LTT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
            LTT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
            configured = true;
        }
        return 0;
    }
    // how many clusters of size S can be resident at once (1 CTA per SM, clusters stay inside a GPC)
    static int max_clusters(int S) {
        static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (cache[S]) return cache[S];
        if (configure()) return -1;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(S * 64, 1, 1);
        cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = SM::TOTAL;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = S;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, STAGES, false, true>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = -1;
        }
        cache[S] = n > 0 ? n : -1;
        return cache[S];
    }
    // CTA-pair launch: clusters of 2, persistent over (M-tile pair, N tile) units
    static int launch_pair(GemmDeviceArgs& a, int cluster_units, cudaStream_t stream) {
        if (int rc = configure()) return rc;
        a.splits = 1;
        a.pair = 1;
        const int mc = max_clusters(2);
        if (mc <= 0) {
            set_error("gemm: no resident 2-CTA clusters for BN=%d", BN);
            return -1;
        }
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * std::min(cluster_units, mc), 1, 1);
        cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = SM::TOTAL;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if (lnf(a)) LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, true, true>, a));
        else LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, true, false>, a));
        LTT_CUDA_OK(cudaGetLastError());
        return 0;
    }
    static int launch(GemmDeviceArgs& a, int ctas, int S, int num_sms, cudaStream_t stream) {
        if (int rc = configure()) return rc;
        a.splits = S;
        if (S == 1) {
            const int grid = ctas < num_sms * GEMM_MINB ? ctas : num_sms * GEMM_MINB;
            if (lnf(a)) LTT_CUDA_OK(launch_k(gemm_tc_kernel<BN, STAGES, false, true>, dim3(grid), dim3(GEMM_THREADS), SM::TOTAL, stream, a));
            else LTT_CUDA_OK(launch_k(gemm_tc_kernel<BN, STAGES, false, false>, dim3(grid), dim3(GEMM_THREADS), SM::TOTAL, stream, a));
        } else {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(ctas * S, 1, 1);
            cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
            cfg.dynamicSmemBytes = SM::TOTAL;
            cfg.stream = stream;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = S;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = pdl_enabled() ? 2 : 1;
            if (lnf(a)) LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, false, true>, a));
            else LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, false, false>, a));
        }
        LTT_CUDA_OK(cudaGetLastError());
        return 0;
    }
};

int gemm_tc_launch(const GemmProblem& p, int num_sms, cudaStream_t stream, int* stats_slots, int* chosen) {
    GemmDeviceArgs a;
    memset(&a, 0, sizeof(a));
    a.B = p.B; a.H = p.H; a.W = p.W; a.N = p.N;
    a.nsrc = p.nsrc;
    pick_tile(p.B, p.H, p.W, &a.tw, &a.th);
    const int tb = BM / (a.tw * a.th);
    a.tiles_x = (p.W + a.tw - 1) / a.tw;
    a.tiles_y = (p.H + a.th - 1) / a.th;
    const int tiles_b = (p.B + tb - 1) / tb;
    const int mtiles = a.tiles_x * a.tiles_y * tiles_b;
    int iters = 0, ktot = 0;
    for (int s = 0; s < p.nsrc; ++s) {
        const GemmSrc& g = p.src[s];
        if (g.channels % BK != 0 || (g.taps != 1 && g.taps != 9) || g.ld % 8 != 0) {
            set_error("gemm: bad source %d (channels %d, taps %d, ld %d)", s, g.channels, g.taps, g.ld);
            return -1;
        }
        a.taps[s] = g.taps;
        a.kchunks[s] = g.channels / BK;
        iters += g.taps * a.kchunks[s];
        ktot += g.taps * g.channels;
        uint64_t dims[4] = {(uint64_t)g.channels, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
        uint64_t str[3] = {(uint64_t)g.ld, (uint64_t)g.ld * p.W, (uint64_t)g.ld * p.W * p.H};
        uint32_t box[4] = {(uint32_t)BK, (uint32_t)a.tw, (uint32_t)a.th, (uint32_t)tb};
        int rc = make_tmap_f16(&a.amap[s], g.ptr, 4, dims, str, box);
        if (rc) return rc;
    }
    if (ktot != p.Ktot || p.N % 8 != 0) {
        set_error("gemm: K mismatch (%d vs %d) or N %% 8 (N=%d)", ktot, p.Ktot, p.N);
        return -1;
    }
    a.iters_total = iters;
    a.epi = p.epi;
    a.kind = epilogue_kind(p.epi);
    static const bool no_prefetch = getenv("LTT_NO_WPREFETCH") != nullptr;
    a.w_static = p.w_static && !no_prefetch;

    // Tile width BN and split-K cluster size S from a small cycle model (constants fitted to the measurements in
    // profiles/): per k-iteration a CTA is bound by the slower of the tensor pipe (2*BN cycles for 128 x BN x 64) and
    // its share of L2->SM bandwidth (~75 B/cycle/SM); the epilogue costs ~600 + 6*BN cycles per tile; a cluster
    // reduction moves 128*BN*4 B per CTA through DSMEM (~20 B/cycle) plus the dump; launch + prologue ~3000 cycles.
    const bool geglu = p.epi.act == ACT_GEGLU;
    if (p.epi.act == ACT_QUICKGELU && (p.epi.res || p.epi.has_gate || p.epi.rowvec || p.epi.out_dtype != DT_F16 || p.epi.out_mode != OUT_ROWMAJOR)) {
        set_error("gemm: the quick-GELU epilogue takes bias -> activation -> fp16 only");
        return -1;
    }
    if (geglu && p.N % 128 != 0) {
        set_error("gemm: GEGLU needs N %% 128 == 0 (N=%d)", p.N);
        return -1;
    }
    // CTA-pair mode (cta_group::2): per k-iteration a lone CTA moves (16 KB A + 128*bn B) through shared memory twice
    // (TMA write + tensor-core read) against 128 B/cycle, which is what bounds the K-deep GEMMs; in a pair each CTA
    // stages only half of the B tile.  Used for main-loop-bound shapes (>= 10 k-iterations, >= 2 M tiles).
    static const int pair_mode = getenv("LTT_GEMM_PAIR") ? atoi(getenv("LTT_GEMM_PAIR")) : 1;
    static const int kBN[4] = {64, 128, 160, 256};
    int BN = 128, S = 1;
    bool use_pair = false;
    double best = 1e30;
    static const int force_bn = getenv("LTT_GEMM_BN") ? atoi(getenv("LTT_GEMM_BN")) : 0;     // experiments only
    const bool forced = p.force_bn != 0;           // autotuner: exactly this tiling or GEMM_ILLEGAL_TILING
    for (int bi = 0; bi < 4; ++bi) {
        const int bn = kBN[bi];
        if (geglu && bn % 128) continue;
        if (LTT_LITE && bn > 128) continue;
        if (forced && bn != p.force_bn) continue;
        if (!forced && force_bn && bn != force_bn && !(geglu && force_bn % 128)) continue;
        const int nt = (p.N + bn - 1) / bn;
        const int ctas_c = mtiles * nt;
        const bool want_stats = p.epi.stats_out != nullptr;
        if (want_stats && nt * EPI_WGS > p.epi.stats_ld) continue;      // row-statistics slots of this tiling must fit
        const double it_cycles = std::max(2.0 * bn, (16384.0 + 128.0 * bn) / 75.0);
        const double epi_cycles = 600.0 + 6.0 * bn;
        if (forced ? (p.force_pair && mtiles >= 2 && bn >= 128) : (pair_mode && mtiles >= 2 && (iters >= 10 || pair_mode == 2) && bn >= 128)) {
            int mc = 0;
            switch (bn) {
                case 128: mc = Variant<128, ST128>::max_clusters(2); break;
                case 160: mc = Variant<160, ST160>::max_clusters(2); break;
                default: mc = Variant<256, ST256>::max_clusters(2); break;
            }
            const int cu = ((mtiles + 1) / 2) * nt;
            // measured: the pair only pays once every CTA walks several tiles (cluster sync + remote barriers cost ~1.5 us)
            if (mc > 0 && (cu >= 2 * mc || pair_mode == 2 || forced)) {
                const int waves = (cu + mc - 1) / mc;
                const double it_pair = 1.1 * std::max(2.0 * bn, 256.0 + bn);
                const double per_unit = iters * it_pair;
                const double t = 3500.0 + (waves > 1 ? waves * std::max(per_unit, epi_cycles) + std::min(per_unit, epi_cycles)
                                                     : per_unit + epi_cycles);
                if (t < best || (pair_mode == 2 && !use_pair) || forced) {
                    best = (pair_mode == 2 || forced) ? 0.0 : t;      // 2: force pair mode wherever it is legal (experiments)
                    BN = bn;
                    S = 1;
                    use_pair = true;
                }
            }
        }
        static const int max_split = getenv("LTT_GEMM_MAXSPLIT") ? atoi(getenv("LTT_GEMM_MAXSPLIT")) : 8;     // experiments only
        for (int sp = 1; sp <= max_split; ++sp) {
            if (forced && (p.force_pair || sp != p.force_splits)) continue;
            int resident = num_sms;
            if (sp > 1) {
                if (p.epi.act == ACT_QUICKGELU) break;      // the cluster-reduction path runs the generic epilogue, which has no quick-GELU
                if (sp * 2 > iters) break;
                const int per = (iters + sp - 1) / sp;
                if ((iters + per - 1) / per != sp) continue;      // every rank must own at least one iteration
                int mc = 0;
                switch (bn) {
                    case 64: mc = Variant<64, ST64>::max_clusters(sp); break;
                    case 128: mc = Variant<128, ST128>::max_clusters(sp); break;
                    case 160: mc = Variant<160, ST160>::max_clusters(sp); break;
                    default: mc = Variant<256, ST256>::max_clusters(sp); break;
                }
                if (mc < ctas_c) continue;                         // all clusters of the launch must be co-resident
                if (want_stats && nt * sp * EPI_WGS > p.epi.stats_ld) continue;
                resident = mc * sp;
            }
            const int units = ctas_c * sp;
            const int waves = (units + resident - 1) / resident;
            // several units per CTA overlap epilogue and main loop (double-buffered accumulator)
            const double per_unit = (double)((iters + sp - 1) / sp) * it_cycles;
            double t = 3000.0 + (waves > 1 ? waves * std::max(per_unit, epi_cycles) + std::min(per_unit, epi_cycles)
                                           : per_unit + epi_cycles);
            if (sp > 1) t += 30.0 * bn + 800.0;
            if (t < best) {
                best = t;
                BN = bn;
                S = sp;
                use_pair = false;
            }
        }
    }
    const int ntiles = (p.N + BN - 1) / BN;
    const int ctas = mtiles * ntiles;
    if (best >= 1e30) {
        if (forced) return GEMM_ILLEGAL_TILING;
        set_error("gemm: no tiling fits (N=%d, stats_ld=%d)", p.N, p.epi.stats_ld);
        return -1;
    }
    if (p.epi.ln_stats && (!p.epi.ln_s || p.epi.ln_K <= 0 || (p.epi.ln_slots <= 0 && p.epi.ln_slots != -1))) {
        set_error("gemm: incomplete LayerNorm-fold operands");
        return -1;
    }
    if (stats_slots) *stats_slots = ntiles * S * EPI_WGS;
    if (chosen) {
        chosen[0] = BN; chosen[1] = S; chosen[2] = use_pair ? 1 : 0;
    }
    a.mtiles = mtiles;
    a.ntiles = ntiles;
    {
        uint64_t dims[2] = {(uint64_t)p.Ktot, (uint64_t)p.N};
        uint64_t str[1] = {(uint64_t)p.Ktot};
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)(use_pair ? BN / 2 : BN)};
        int rc = make_tmap_f16(&a.bmap, p.w, 2, dims, str, box);
        if (rc) return rc;
    }
    if (use_pair) {
        const int cu = ((mtiles + 1) / 2) * ntiles;
        switch (BN) {
            case 128: return Variant<128, ST128>::launch_pair(a, cu, stream);
            case 160: return Variant<160, ST160>::launch_pair(a, cu, stream);
            default: return Variant<256, ST256>::launch_pair(a, cu, stream);
        }
    }
    switch (BN) {
        case 64: return Variant<64, ST64>::launch(a, ctas, S, num_sms, stream);
        case 128: return Variant<128, ST128>::launch(a, ctas, S, num_sms, stream);
        case 160: return Variant<160, ST160>::launch(a, ctas, S, num_sms, stream);
        case 256: return Variant<256, ST256>::launch(a, ctas, S, num_sms, stream);
    }
    set_error("gemm: no kernel variant for BN=%d", BN);
    return -1;
}

}  // namespace ltt
