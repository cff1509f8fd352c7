// CUDA-core kernels of the UNet path (everything that is not a dense contraction): GroupNorm statistics / apply
// (+SiLU), LayerNorm, first / last 3x3 convs with 4 channels, nearest upsample, stride-2 patch gather, timestep and
// Fourier box embeddings, the relation-aware box pooling / scatter, the 30x10 relation attention (warp shuffles)
// and the fused CFG + PLMS update.  All HBM/L2-bound elementwise or reduction work; coalesced on the channel axis.
//
// Reference semantics (paths under /root/reference/GLIGEN/ldm): GroupNorm32 util.py:211-229 (fp32 stats, eps 1e-5),
// Normalize attention.py:78-79 (eps 1e-6), nn.LayerNorm eps 1e-5, timestep_embedding util.py:161-181,
// FourierEmbedder util.py:12-26, PositionNet text_grounding_net.py:26-43, RelationCrossAttention
// attention.py:315-359, p_sample_plms models/diffusion/plms.py:110-163.
#include <cstdio>
#include <cstdlib>

#include "ltt_ops.h"
#include "ltt_ptx.cuh"

namespace ltt {

__device__ __forceinline__ float r16f(float x) { return __half2float(__float2half_rn(x)); }
// x * sigmoid(x) with ex2.approx / rcp.approx (2 MUFU + 3 FMA-pipe ops; the IEEE division cost ~15 instructions per
// element and made GroupNorm+SiLU instruction bound).  Relative error ~2e-7, far below the fp16 rounding that follows.
__device__ __forceinline__ float siluf(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// stats[b][g] = {sum, sumsq} in double (zeroed by the caller).  Input = channel concat of up to two NHWC tensors.
__global__ void gn_stats_kernel(const __half* __restrict__ x0, int c0, int ld0, const __half* __restrict__ x1, int c1,
                                int ld1, int HW, int cpg, int strip, double* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double sh[64];
    const int b = blockIdx.y, C = c0 + c1, P = C >> 1;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int p0 = blockIdx.x * strip, p1 = min(HW, p0 + strip);
    for (int cp = threadIdx.x; cp < P; cp += blockDim.x) {
        const int c = cp * 2;
        const __half* src;
        int ld, cc;
        if (c < c0) { src = x0; ld = ld0; cc = c; } else { src = x1; ld = ld1; cc = c - c0; }
        float s = 0.f, ss = 0.f;
        for (int p = p0; p < p1; ++p) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(src + ((size_t)b * HW + p) * ld + cc));
            s += v.x + v.y;
            ss += v.x * v.x + v.y * v.y;
        }
        const int g = c / cpg;
        atomicAdd(&sh[2 * g], (double)s);
        atomicAdd(&sh[2 * g + 1], (double)ss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64; i += blockDim.x) atomicAdd(&stats[(size_t)b * 64 + i], sh[i]);
}

// out[b, p, c] = act(r16((x - mean) * rstd * gamma + beta)); fp16 NHWC out with pitch C.  Each thread handles 8
// consecutive channels (one 16-byte vector); mean / rstd of every (batch, group) are decoded once per block.
__global__ void gn_apply_kernel(const __half* __restrict__ x0, int c0, int ld0, const __half* __restrict__ x1, int c1,
                                int ld1, int B, int HW, int cpg, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                __half* __restrict__ out, size_t total_vecs) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float2 mr[];    // [B*32] (mean, rstd)
    for (int i = threadIdx.x; i < B * 32; i += blockDim.x) {
        const double n = (double)cpg * HW;
        const double mean = stats[2 * i] / n;
        double var = stats[2 * i + 1] / n - mean * mean;
        if (var < 0) var = 0;
        mr[i] = make_float2((float)mean, rsqrtf((float)var + eps));
    }
    __syncthreads();
    const int C = c0 + c1, V = C >> 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % V);
        const size_t row = i / V;            // b*HW + p
        const int b = (int)(row / HW);
        const int c = cv * 8;
        const __half* src;
        int ld, cc;
        if (c < c0) { src = x0; ld = ld0; cc = c; } else { src = x1; ld = ld1; cc = c - c0; }
        const uint4 u = *reinterpret_cast<const uint4*>(src + row * ld + cc);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        float g[8], bt[8];
        {
            const float4 ga = *reinterpret_cast<const float4*>(gamma + c), gb = *reinterpret_cast<const float4*>(gamma + c + 4);
            const float4 ba = *reinterpret_cast<const float4*>(beta + c), bb = *reinterpret_cast<const float4*>(beta + c + 4);
            g[0] = ga.x; g[1] = ga.y; g[2] = ga.z; g[3] = ga.w; g[4] = gb.x; g[5] = gb.y; g[6] = gb.z; g[7] = gb.w;
            bt[0] = ba.x; bt[1] = ba.y; bt[2] = ba.z; bt[3] = ba.w; bt[4] = bb.x; bt[5] = bb.y; bt[6] = bb.z; bt[7] = bb.w;
        }
        __half2 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 m = mr[b * 32 + (c + 2 * k) / cpg];     // cpg is even: a channel pair never straddles groups
            const float2 v = __half22float2(h[k]);
            float a = (v.x - m.x) * m.y * g[2 * k] + bt[2 * k];
            float d = (v.y - m.x) * m.y * g[2 * k + 1] + bt[2 * k + 1];
            if (silu != 2) {      // 2: swish on the fp32 normalised value (VAE: GroupNorm output stays fp32 under autocast)
                a = r16f(a);
                d = r16f(d);
            }
            if (silu) {
                a = siluf(a);
                d = siluf(d);
            }
            o[k] = __floats2half2_rn(a, d);
        }
        *reinterpret_cast<uint4*>(out + row * C + c) = *reinterpret_cast<uint4*>(o);
    }
}

int gn_stats_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                    double* stats, cudaStream_t st) {
    const int C = c0 + c1, cpg = C / groups;
    if (groups != 32 || C % groups || (cpg & 1) || (c0 & 1)) {
        set_error("groupnorm: unsupported C=%d groups=%d", C, groups);
        return -1;
    }
    LTT_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)B * 64 * sizeof(double), st));
    const int strip = HW >= 4096 ? 32 : (HW >= 1024 ? 16 : 8);
    dim3 grid((HW + strip - 1) / strip, B);
    LTT_CUDA_OK(launch_k(gn_stats_kernel, dim3(grid), dim3(256), 0, st, x0, c0, ld0, x1, c1, ld1, HW, cpg, strip, stats));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int gn_apply_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                    const double* stats, const float* gamma, const float* beta, float eps, int silu, __half* out,
                    cudaStream_t st) {
    const int C = c0 + c1, cpg = C / groups;
    if (C % 8 || c0 % 8 || ld0 % 8 || (c1 && ld1 % 8) || B > 64) {
        set_error("groupnorm: unsupported geometry C=%d c0=%d B=%d", C, c0, B);
        return -1;
    }
    const size_t total = (size_t)B * HW * (C / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
    LTT_CUDA_OK(launch_k(gn_apply_kernel, dim3(blocks), dim3(256), (size_t)B * 32 * sizeof(float2), st, x0, c0, ld0, x1, c1, ld1, B, HW, cpg, stats, gamma, beta,
                                                                         eps, silu, out, total));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- single-kernel GroupNorm.  One thread-block cluster per (batch element, set of G consecutive groups); the
// GN_S CTAs of a cluster split the pixels.  Pass 1 streams the CTA's slab (pixels x G*cpg channels) from global memory
// once -- staging it in shared memory when it fits -- and reduces (sum, sum of squares): fp32 per thread over its
// pixel column, then double across threads, then double across the cluster through distributed shared memory in fixed
// rank order (deterministic).  Pass 2 normalises from the staged copy.  One read + one write of the tensor, no
// statistics buffer, no memset node between the producer GEMM and the consumer (the PDL chain stays intact).
constexpr int GN_S = 8;          // CTAs per cluster (pixel split); 16 (non-portable cluster size) for small batches, see the launcher
constexpr int GN_SMAX = 16;
constexpr int GN_MAXG = 32;      // groups per cluster the kernel supports (the launcher picks <= its own limit)

template <int GN_THREADS>
__global__ void __launch_bounds__(GN_THREADS) gn_fused_kernel(
    const __half* __restrict__ x0, int c0, int ld0, const __half* __restrict__ x1, int ld1, int C, int HW, int cpg, int G,
    int V, int R, int px_per_cta, int stage, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    int silu, __half* __restrict__ out) {
    const int S = gridDim.x;                           // CTAs of this cluster (8 or 16)
    pdl_launch_dependents();
    extern __shared__ uint4 gn_slab[];                 // [px_per_cta][V] when stage
    __shared__ float part[GN_THREADS * 8];             // per-thread (sum, sumsq) of its 4 channel pairs
    __shared__ float2 sub2[GN_THREADS];                // second-level partial sums
    __shared__ double colsum[GN_THREADS * 2];          // [V*4 pair columns][2]  (V <= GN_THREADS/4 by host check)
    __shared__ double cta_stats[2 * GN_MAXG];          // this CTA's (sum, sumsq) per group -- read by the whole cluster
    __shared__ double all_stats[2 * GN_MAXG * GN_SMAX];   // every rank's cta_stats, gathered through DSMEM
    __shared__ float2 mr[GN_MAXG];
    const int tid = threadIdx.x;
    const int rank = blockIdx.x;                       // cluster dims (GN_S,1,1), gridDim.x == GN_S
    const int b = blockIdx.z;
    const int cbase = blockIdx.y * G * cpg;
    const bool active = tid < V * R;
    const int v = tid % V, r = tid / V;
    const int c = cbase + v * 8;
    const __half* src = c < c0 ? x0 + c : x1 + (c - c0);
    const int ld = c < c0 ? ld0 : ld1;
    const int p0 = rank * px_per_cta, p1 = min(HW, p0 + px_per_cta);
    // step-invariant operands first (they do not depend on the predecessor grid)
    float g[8], bt[8];
    if (active) {
        const float4 ga = *reinterpret_cast<const float4*>(gamma + c), gb = *reinterpret_cast<const float4*>(gamma + c + 4);
        const float4 ba = *reinterpret_cast<const float4*>(beta + c), bb = *reinterpret_cast<const float4*>(beta + c + 4);
        g[0] = ga.x; g[1] = ga.y; g[2] = ga.z; g[3] = ga.w; g[4] = gb.x; g[5] = gb.y; g[6] = gb.z; g[7] = gb.w;
        bt[0] = ba.x; bt[1] = ba.y; bt[2] = ba.z; bt[3] = ba.w; bt[4] = bb.x; bt[5] = bb.y; bt[6] = bb.z; bt[7] = bb.w;
    }
    pdl_wait();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int UNR = 8;      // independent 16-byte loads in flight per thread (the pass is latency bound otherwise)
    if (active) {
        for (int pb = p0 + r; pb < p1; pb += R * UNR) {
            uint4 u[UNR];
#pragma unroll
            for (int i = 0; i < UNR; ++i) {
                const int p = pb + i * R;
                if (p < p1) u[i] = *reinterpret_cast<const uint4*>(src + ((size_t)b * HW + p) * ld);
            }
#pragma unroll
            for (int i = 0; i < UNR; ++i) {
                const int p = pb + i * R;
                if (p < p1) {
                    if (stage) gn_slab[(p - p0) * V + v] = u[i];
                    const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = __half22float2(h[k]);
                        s[k] += f.x + f.y;
                        ss[k] += f.x * f.x + f.y * f.y;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            part[tid * 8 + 2 * k] = s[k];
            part[tid * 8 + 2 * k + 1] = ss[k];
        }
    }
    __syncthreads();
    const int PC = V * 4;                              // channel-pair columns of the slab
    // two-level reduction over the R pixel rows of every pair column (fixed order: deterministic)
    int SUB = 1;
    while (SUB < 8 && PC * SUB * 2 <= GN_THREADS) SUB *= 2;
    if (tid < PC * SUB) {
        const int pc = tid / SUB, sub = tid - pc * SUB;
        const int vv = pc >> 2, k = pc & 3;
        float a = 0.f, q = 0.f;
        for (int rr = sub; rr < R; rr += SUB) {
            a += part[(rr * V + vv) * 8 + 2 * k];
            q += part[(rr * V + vv) * 8 + 2 * k + 1];
        }
        sub2[tid] = make_float2(a, q);
    }
    __syncthreads();
    if (tid < PC) {
        double a = 0.0, q = 0.0;
        for (int i = 0; i < SUB; ++i) {
            const float2 f = sub2[tid * SUB + i];
            a += (double)f.x;
            q += (double)f.y;
        }
        colsum[2 * tid] = a;
        colsum[2 * tid + 1] = q;
    }
    __syncthreads();
    if (tid < 2 * G) {
        const int gi = tid >> 1, which = tid & 1, npc = cpg >> 1;
        double a = 0.0;
        for (int j = 0; j < npc; ++j) a += colsum[2 * (gi * npc + j) + which];
        cta_stats[tid] = a;
    }
    cluster_sync_all();                                // every CTA's cta_stats is complete and visible
    for (int i = tid; i < 2 * G * S; i += GN_THREADS) {     // remote loads, all in flight together
        const int rk = i / (2 * G), e = i % (2 * G);
        all_stats[rk * 2 * GN_MAXG + e] = dsmem_ld_f64(dsmem_map(smem_u32(&cta_stats[e]), rk));
    }
    cluster_arrive();                                  // done reading peers; matching wait at the end
    __syncthreads();
    if (tid < G) {
        double sm = 0.0, sq = 0.0;
        for (int rk = 0; rk < S; ++rk) {               // fixed rank order: deterministic
            sm += all_stats[rk * 2 * GN_MAXG + 2 * tid];
            sq += all_stats[rk * 2 * GN_MAXG + 2 * tid + 1];
        }
        const double n = (double)cpg * HW;
        const double mean = sm / n;
        double var = sq / n - mean * mean;
        if (var < 0) var = 0;
        mr[tid] = make_float2((float)mean, rsqrtf((float)var + eps));
    }
    __syncthreads();
    if (active) {
        // y = x * a + b with a = rstd * gamma, b = beta - mean * a (the formulation of PyTorch's own GroupNorm kernel)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 m = mr[(v * 8 + 2 * k) / cpg];   // cpg is even: a pair never straddles groups
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float a = m.y * g[2 * k + e];
                bt[2 * k + e] = bt[2 * k + e] - m.x * a;
                g[2 * k + e] = a;
            }
        }
        constexpr int UN2 = 4;
        for (int pb = p0 + r; pb < p1; pb += R * UN2) {
            uint4 u[UN2];
#pragma unroll
            for (int i = 0; i < UN2; ++i) {
                const int p = pb + i * R;
                if (p < p1)
                    u[i] = stage ? gn_slab[(p - p0) * V + v] : *reinterpret_cast<const uint4*>(src + ((size_t)b * HW + p) * ld);
            }
#pragma unroll
            for (int i = 0; i < UN2; ++i) {
                const int p = pb + i * R;
                if (p >= p1) continue;
                const __half2* h = reinterpret_cast<const __half2*>(&u[i]);
                __half2 o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(h[k]);
                    const float y0 = fmaf(f.x, g[2 * k], bt[2 * k]), y1 = fmaf(f.y, g[2 * k + 1], bt[2 * k + 1]);
                    __half2 y;
                    if (silu == 2) {      // swish on the fp32 value, one rounding (VAE blocks under autocast)
                        y = __floats2half2_rn(siluf(y0), siluf(y1));
                    } else {
                        y = __floats2half2_rn(y0, y1);
                        if (silu) {       // GroupNorm32 casts back to fp16 before the SiLU (UNet ResBlock)
                            const float2 yf = __half22float2(y);
                            y = __floats2half2_rn(siluf(yf.x), siluf(yf.y));
                        }
                    }
                    o[k] = y;
                }
                *reinterpret_cast<uint4*>(out + ((size_t)b * HW + p) * C + c) = *reinterpret_cast<uint4*>(o);
            }
        }
    }
    cluster_wait();                                    // no CTA exits while a peer may still read its cta_stats
}

// GroupNorm(32 groups) (+SiLU) of the channel concat of up to two NHWC fp16 tensors -> NHWC fp16 [B, HW, C].
// Returns 1 when the geometry is not supported by the fused kernel (caller falls back to stats + apply).
int groupnorm_fused_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                           const float* gamma, const float* beta, float eps, int silu, __half* out, cudaStream_t st) {
    const int C = c0 + c1;
    if (groups != 32 || C % groups) return 1;
    const int cpg = C / groups;
    if ((cpg & 1) || c0 % 8 || ld0 % 8 || (c1 && ld1 % 8)) return 1;
    // groups per cluster: the slab must be whole 16-byte vectors; fewer groups per cluster = more CTAs for small batches
    int G = 0;
    // groups per cluster: more groups = wider contiguous row segments per CTA (cpg * G channels), fewer clusters.  The
    // limit only matters once there are >= 128 CTAs anyway (B >= 16 per launch): measured 4 -> 8: 16x4096x320 40.2 -> 36.3 us,
    // 16x256x1280 20.0 -> 13.3 us, 128x1024x640 215 -> 159 us, unchanged at B = 2; 16 / 32 are mixed (profiles/r02_gn_ab.txt)
    static const int maxg = getenv("LTT_GN_MAXG") ? atoi(getenv("LTT_GN_MAXG")) : 8;
    for (int g = 1; g <= std::min(maxg, GN_MAXG); g *= 2)
        if ((g * cpg) % 8 == 0 && g * cpg / 8 <= 64 && (G == 0 || GN_S * B * (32 / g) >= 128)) G = g;
    if (G == 0) return 1;
    const int V = G * cpg / 8;
    // 16 CTAs per cluster (non-portable size) halve both passes over the slab, but the wider cluster barrier / gang launch
    // costs more than that saves: measured at B = 2 (one image): 4096 x 320: 9.9 -> 11.8 us, 1024 x 640: 6.7 -> 9.6 us, only
    // the 960-channel concat gains (23.2 -> 19.0 us); 270.9 -> 273.6 ms per image.  Off by default (LTT_GN_S16=1: A/B).
    static const int s16 = getenv("LTT_GN_S16") ? atoi(getenv("LTT_GN_S16")) : 0;
    // (smaller clusters for the small feature maps were measured too -- 2 x 256 x 1280: S = 8 5.5 us, 4: 5.9, 2: 7.8, 1: 11.3 us;
    // 2 x 1024 x 640: 8: 6.9, 4: 8.7, 2: 11.5 us -- the cluster barrier is cheaper than the longer per-CTA pixel loop)
    const int S = (s16 && HW >= 1024 && 16 * B * (32 / G) <= 296) ? 16 : GN_S;
    const int px = (HW + S - 1) / S;
    const size_t slab = (size_t)px * V * 16;
    // big slabs (level-0 concat inputs) own their SM anyway: 16 warps hide the load / MUFU latency better than 8
    const int threads = slab > 96 * 1024 ? 512 : 256;
    if (V > threads / 4) return 1;
    const int R = threads / V;
    const int stage = slab <= 160 * 1024;
    static bool configured = false;
    if (!configured) {
        LTT_CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        LTT_CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        LTT_CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        LTT_CUDA_OK(cudaFuncSetAttribute(gn_fused_kernel<512>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(S, 32 / G, B);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = stage ? slab : 0;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (threads == 512)
        LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gn_fused_kernel<512>, x0, c0, ld0, x1, ld1, C, HW, cpg, G, V, R, px, stage, gamma, beta, eps,
                                       silu, out));
    else
    LTT_CUDA_OK(cudaLaunchKernelEx(&cfg, gn_fused_kernel<256>, x0, c0, ld0, x1, ld1, C, HW, cpg, G, V, R, px, stage, gamma, beta, eps,
                                   silu, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, 16-byte vector loads / stores (8 halves or 2 x 4 floats per lane per step); fp32 statistics
// (two pass in registers); writes fp16 and/or fp32.
template <typename T>
__device__ __forceinline__ void ln_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ln_load8<__half>(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void ln_load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, __half* __restrict__ out16,
                                 float* __restrict__ out32, float2* __restrict__ stats_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const T* xr = x + (size_t)row * C;
    const int nvec = C >> 3;           // host checks C % 8 == 0, C <= 1280 (<= 5 vectors per lane)
    float v[5][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int j = i * 32 + lane;
        if (j < nvec) {
            ln_load8<T>(xr + j * 8, v[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[i][k];
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int j = i * 32 + lane;
        if (j < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float d = v[i][k] - mean;
                ss += d * d;
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / C + eps);
    if (stats_out) {                   // statistics only: consumers normalise on the fly (relation pool / scatter)
        if (lane == 0) stats_out[row] = make_float2(mean, rstd);
        return;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int j = i * 32 + lane;
        if (j < nvec) {
            const int c = j * 8;
            float g[8], bt[8], y[8];
            ln_load8<float>(gamma + c, g);
            ln_load8<float>(beta + c, bt);
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = (v[i][k] - mean) * rstd * g[k] + bt[k];
            if (out16) {
                __half2 h[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(y[2 * k], y[2 * k + 1]);
                *reinterpret_cast<uint4*>(out16 + (size_t)row * C + c) = *reinterpret_cast<uint4*>(h);
            }
            if (out32) {
                float4* o = reinterpret_cast<float4*>(out32 + (size_t)row * C + c);
                o[0] = make_float4(y[0], y[1], y[2], y[3]);
                o[1] = make_float4(y[4], y[5], y[6], y[7]);
            }
        }
    }
}

// Many rows (batch >= 8 images): the one-row-per-warp kernel re-reads gamma / beta (64 B of L1 traffic per 16 B of data) for
// every row and keeps one row per warp in flight; here a warp holds its lanes' gamma / beta in registers and walks rows with a
// grid stride, D rows ahead of the one being reduced held as RAW 16-byte vectors (fp16 rows are 640 B at C = 320: the bytes
// in flight per SM, not the arithmetic, set the rate).  NV = vectors per lane (C <= 256*NV).
template <typename T>
struct LnRaw;
template <>
struct LnRaw<__half> {
    uint4 a;
    __device__ __forceinline__ void load(const __half* p) { a = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
};
template <>
struct LnRaw<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *reinterpret_cast<const float4*>(p + 4);
    }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

template <typename T, int NV, int D>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const T* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, __half* __restrict__ out16,
                                                              float* __restrict__ out32) {
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int nvec = C >> 3;
    float g[NV][8], bt[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = i * 32 + lane;
        if (j < nvec) {
            ln_load8<float>(gamma + j * 8, g[i]);
            ln_load8<float>(beta + j * 8, bt[i]);
        }
    }
    pdl_wait();
    const int wstride = gridDim.x * (blockDim.x >> 5);
    const int row0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    LnRaw<T> ring[D][NV];
    auto load_row = [&](int r, LnRaw<T> (&dst)[NV]) {
        const T* xr = x + (size_t)r * C;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = i * 32 + lane;
            if (j < nvec) dst[i].load(xr + j * 8);
        }
    };
#pragma unroll
    for (int d = 0; d < D; ++d)
        if (row0 + d * wstride < M) load_row(row0 + d * wstride, ring[d]);
    for (int base = row0; base < M; base += D * wstride) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int row = base + d * wstride;
            if (row >= M) break;
            float v[NV][8];
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i * 32 + lane < nvec) ring[d][i].unpack(v[i]);
            if (row + D * wstride < M) load_row(row + D * wstride, ring[d]);      // in flight during D rows of arithmetic
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i * 32 + lane < nvec) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) s += v[i][k];
                }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / C;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i * 32 + lane < nvec) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float dd = v[i][k] - mean;
                        ss += dd * dd;
                    }
                }
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = rsqrtf(ss / C + eps);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = i * 32 + lane;
                if (j < nvec) {
                    float y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) y[k] = (v[i][k] - mean) * rstd * g[i][k] + bt[i][k];     // same expression as layernorm_kernel
                    if (out16) {
                        __half2 h[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(y[2 * k], y[2 * k + 1]);
                        *reinterpret_cast<uint4*>(out16 + (size_t)row * C + j * 8) = *reinterpret_cast<uint4*>(h);
                    }
                    if (out32) {
                        float4* o = reinterpret_cast<float4*>(out32 + (size_t)row * C + j * 8);
                        o[0] = make_float4(y[0], y[1], y[2], y[3]);
                        o[1] = make_float4(y[4], y[5], y[6], y[7]);
                    }
                }
            }
        }
    }
}

template <typename T, int NV>
static int layernorm_rows_go(const T* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out16, float* out32,
                             cudaStream_t st) {
    static int resident = 0;                   // CTAs per SM of this instantiation: the grid is exactly one resident wave
    if (!resident) {
        int n = 0;
        LTT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, layernorm_rows_kernel<T, NV, (sizeof(T) == 2 ? 4 : 2)>, 256, 0));
        resident = n > 0 ? n : 1;
    }
    LTT_CUDA_OK(launch_k(layernorm_rows_kernel<T, NV, (sizeof(T) == 2 ? 4 : 2)>, dim3(148 * resident), dim3(256), 0, st, x, M, C, gamma, beta, eps, out16, out32));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T>
static int layernorm_rows_launch(const T* x, int M, int C, const float* gamma, const float* beta, float eps, __half* out16,
                                 float* out32, cudaStream_t st) {
    const int nvec = C >> 3;
    if (nvec <= 32) return layernorm_rows_go<T, 1>(x, M, C, gamma, beta, eps, out16, out32, st);
    if (nvec <= 64) return layernorm_rows_go<T, 2>(x, M, C, gamma, beta, eps, out16, out32, st);
    return layernorm_rows_go<T, 3>(x, M, C, gamma, beta, eps, out16, out32, st);
}

int layernorm_launch(const void* x, int x_dtype, int M, int C, const float* gamma, const float* beta, float eps,
                     __half* out16, float* out32, cudaStream_t st, float2* stats_out) {
    if (C % 8 || C > 1280) {
        set_error("layernorm: unsupported C=%d", C);
        return -1;
    }
    // rows >> resident warps and C <= 768: the grid-stride variant (LTT_LN_ROWS=0: always one row per warp, A/B)
    static const int rows_variant = getenv("LTT_LN_ROWS") ? atoi(getenv("LTT_LN_ROWS")) : 1;
    if (rows_variant && !stats_out && C <= 512 && M >= 16384) {      // C = 640 (2.5 vectors per lane) measured slower: 3.8 -> 3.4 TB/s
        if (x_dtype == DT_F16) return layernorm_rows_launch((const __half*)x, M, C, gamma, beta, eps, out16, out32, st);
        return layernorm_rows_launch((const float*)x, M, C, gamma, beta, eps, out16, out32, st);
    }
    const int wpb = 8;
    const int blocks = (M + wpb - 1) / wpb;
    if (x_dtype == DT_F16)
        LTT_CUDA_OK(launch_k(layernorm_kernel<__half>, dim3(blocks), dim3(wpb * 32), 0, st, (const __half*)x, M, C, gamma, beta, eps, out16, out32, stats_out));
    else
        LTT_CUDA_OK(launch_k(layernorm_kernel<float>, dim3(blocks), dim3(wpb * 32), 0, st, (const float*)x, M, C, gamma, beta, eps, out16, out32, stats_out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ first / last conv
// input_blocks.0.0: Conv2d(4 -> Cout, 3x3, pad 1) on NCHW fp32 x; inputs and weights rounded to fp16 (autocast),
// fp32 accumulate, NHWC fp16 out.  w: [Cout, Cin, 3, 3] fp32.
constexpr int CIN_PIX = 32;   // pixels per block
// One block = CIN_PIX consecutive pixels x all output channels (one thread per channel).  The fp16-rounded weights sit
// in shared memory as [K][Cout + 2] (the +2 keeps the transposing fill free of bank conflicts), the input patches as
// [K][CIN_PIX] so a thread reads four pixels per 16-byte shared load.
__global__ void conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               int B, int Cin, int H, int W, int Cout, __half* __restrict__ out) {
    pdl_launch_dependents();
    extern __shared__ float smem_ci[];
    const int K = Cin * 9, WS = Cout + 2;
    float* patch = smem_ci;                                        // [K][CIN_PIX]
    __half* wt = reinterpret_cast<__half*>(patch + CIN_PIX * K);   // [K][WS], fp16-rounded weights
    const int pix0 = blockIdx.x * CIN_PIX, total = B * H * W;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {     // weights are static: filled before the PDL wait
        const int co = i / K, k = i - co * K;
        wt[k * WS + co] = __float2half_rn(w[i]);
    }
    pdl_wait();
    for (int i = threadIdx.x; i < CIN_PIX * K; i += blockDim.x) {
        const int k = i / CIN_PIX, p = i - k * CIN_PIX, pix = pix0 + p;
        float v = 0.f;
        if (pix < total) {
            const int b = pix / (H * W), rem = pix - b * (H * W), y = rem / W, xx = rem - y * W;
            const int ci = k / 9, t = k - ci * 9, yy = y + t / 3 - 1, xs = xx + t % 3 - 1;
            if (yy >= 0 && yy < H && xs >= 0 && xs < W) v = r16f(x[((size_t)(b * Cin + ci) * H + yy) * W + xs]);
        }
        patch[i] = v;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
        float acc[CIN_PIX];
#pragma unroll
        for (int p = 0; p < CIN_PIX; ++p) acc[p] = 0.f;
        for (int k = 0; k < K; ++k) {
            const float ww = __half2float(wt[k * WS + co]);
            const float4* pr = reinterpret_cast<const float4*>(patch + k * CIN_PIX);
#pragma unroll
            for (int q = 0; q < CIN_PIX / 4; ++q) {
                const float4 pv = pr[q];
                acc[4 * q] += pv.x * ww;
                acc[4 * q + 1] += pv.y * ww;
                acc[4 * q + 2] += pv.z * ww;
                acc[4 * q + 3] += pv.w * ww;
            }
        }
        const float bb = bias[co];
#pragma unroll
        for (int p = 0; p < CIN_PIX; ++p)
            if (pix0 + p < total) out[(size_t)(pix0 + p) * Cout + co] = __float2half_rn(acc[p] + bb);
    }
}

int conv_in_launch(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W, int Cout,
                   __half* out, cudaStream_t st) {
    const int K = Cin * 9, total = B * H * W;
    const size_t smem = (size_t)CIN_PIX * K * sizeof(float) + (size_t)K * (Cout + 2) * sizeof(__half);
    if (smem > 48 * 1024) {
        set_error("conv_in: Cin=%d Cout=%d needs %zu B of shared memory", Cin, Cout, smem);
        return -1;
    }
    const int threads = std::min(320, (Cout + 31) / 32 * 32);
    LTT_CUDA_OK(launch_k(conv_in_kernel, dim3((total + CIN_PIX - 1) / CIN_PIX), dim3(threads), smem, st, x, w, bias, B, Cin, H, W, Cout, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// out.2: Conv2d(C -> Cout(<= 4), 3x3, pad 1) on NHWC fp16 input (already GN+SiLU), w packed [Cout][9][C] fp16,
// NCHW fp32 out (values rounded to fp16 as the autocast reference returns half).  One warp per pixel: lanes split the
// channels in 16-byte vectors, the nine taps' loads are issued back to back; the 23 KB of weights are staged in shared
// memory once per block (filled before the PDL wait -- they are static).
constexpr int COUT_WARPS = 8, COUT_PIX_PER_WARP = 4;
__global__ void __launch_bounds__(COUT_WARPS * 32) conv_out_kernel(const __half* __restrict__ x, const __half* __restrict__ w,
                                                                  const float* __restrict__ bias, int B, int H, int W, int C,
                                                                  int Cout, float* __restrict__ out) {
    pdl_launch_dependents();
    extern __shared__ uint4 smem_co[];                    // [Cout*9*C/8] packed weights
    const int nw = Cout * 9 * C / 8;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) smem_co[i] = reinterpret_cast<const uint4*>(w)[i];
    pdl_wait();
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int V = C >> 3;                                  // 16-byte vectors per pixel
    const int total = B * H * W;
    for (int pp = 0; pp < COUT_PIX_PER_WARP; ++pp) {
        const int pix = (blockIdx.x * COUT_WARPS + warp) * COUT_PIX_PER_WARP + pp;
        if (pix >= total) break;
        const int b = pix / (H * W), rem = pix - b * (H * W), y = rem / W, xx = rem - y * W;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cv = lane; cv < V; cv += 32) {
            uint4 xv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int yy = y + t / 3 - 1, xs = xx + t % 3 - 1;
                const bool in = yy >= 0 && yy < H && xs >= 0 && xs < W;
                xv[t] = in ? *reinterpret_cast<const uint4*>(x + ((size_t)(b * H + yy) * W + xs) * C + cv * 8) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const __half2* xh = reinterpret_cast<const __half2*>(&xv[t]);
#pragma unroll
                for (int co = 0; co < 4; ++co)
                    if (co < Cout) {
                        const uint4 wv = smem_co[(co * 9 + t) * V + cv];
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
            out[((size_t)(b * Cout + lane) * H + y) * W + xx] = r16f(a + bias[lane]);
        }
    }
}

int conv_out_launch(const __half* x, const __half* w, const float* bias, int B, int H, int W, int C, int Cout,
                    float* out, cudaStream_t st) {
    const size_t smem = (size_t)Cout * 9 * C * 2;
    if (Cout > 4 || (C & 7) || smem > 96 * 1024) {
        set_error("conv_out: unsupported Cout=%d C=%d", Cout, C);
        return -1;
    }
    static bool configured = false;
    if (!configured) {
        LTT_CUDA_OK(cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured = true;
    }
    const int per_block = COUT_WARPS * COUT_PIX_PER_WARP, total = B * H * W;
    LTT_CUDA_OK(launch_k(conv_out_kernel, dim3((total + per_block - 1) / per_block), dim3(COUT_WARPS * 32), smem, st, x, w, bias, B, H, W, C, Cout, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ resampling helpers
// nearest 2x upsample, NHWC fp16, 16-byte vectors
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int C8) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * 4 * H * W * C8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        size_t p = i / C8;
        const int ox = (int)(p % (2 * W));
        p /= (2 * W);
        const int oy = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        out[i] = in[((size_t)(b * H + (oy >> 1)) * W + (ox >> 1)) * C8 + c];
    }
}
int upsample2x_launch(const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st) {
    const size_t total = (size_t)B * 4 * H * W * (C / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(upsample2x_kernel, dim3(blocks), dim3(256), 0, st, (const uint4*)in, (uint4*)out, B, H, W, C / 8));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// stride-2 pad-1 3x3 patch gather: out[(b, oy, ox)][tap*C + c] = in[b, 2oy-1+ky, 2ox-1+kx, c] (0 outside)
__global__ void im2col_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int C8) {
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * Ho * Wo * 9 * C8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        size_t p = i / C8;
        const int t = (int)(p % 9);
        p /= 9;
        const int ox = (int)(p % Wo);
        p /= Wo;
        const int oy = (int)(p % Ho);
        const int b = (int)(p / Ho);
        const int yy = 2 * oy - 1 + t / 3, xx = 2 * ox - 1 + t % 3;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = in[((size_t)(b * H + yy) * W + xx) * C8 + c];
        out[i] = v;
    }
}
int im2col_s2_launch(const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2) * 9 * (C / 8);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(im2col_s2_kernel, dim3(blocks), dim3(256), 0, st, (const uint4*)in, (uint4*)out, B, H, W, C / 8));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ embeddings
// timestep_embedding: [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(10000) i / half), fp32 math, fp16 out [B, dim]
__global__ void timestep_embed_kernel(const float* __restrict__ t, int B, int dim, __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int half = dim / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * half; i += gridDim.x * blockDim.x) {
        const int b = i / half, k = i % half;
        const float f = expf(-logf(10000.0f) * (float)k / (float)half);
        const float a = t[b] * f;
        out[(size_t)b * dim + k] = __float2half_rn(cosf(a));
        out[(size_t)b * dim + half + k] = __float2half_rn(sinf(a));
    }
}
int timestep_embed_launch(const float* t, int B, int dim, __half* out, cudaStream_t st) {
    LTT_CUDA_OK(launch_k(timestep_embed_kernel, dim3((B * dim / 2 + 127) / 128), dim3(128), 0, st, t, B, dim, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// PositionNet input rows: [ emb*m + (1-m)*null_txt (768) | fourier(box)*m + (1-m)*null_pos (8 freqs x (sin4|cos4)) ]
// One warp per box; lanes 0..31 cover the (freq, coord) pairs of the Fourier part via shuffles of the 4 coords.
__global__ void posnet_input_kernel(const float* __restrict__ boxes, const float* __restrict__ masks,
                                    const float* __restrict__ emb, const float* __restrict__ null_txt,
                                    const float* __restrict__ null_pos, int rows, int in_dim, int nfreq,
                                    __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float m = masks[row];
    const int width = in_dim + nfreq * 8;
    __half* o = out + (size_t)row * width;
    for (int c = lane; c < in_dim; c += 32) o[c] = __float2half_rn(emb[(size_t)row * in_dim + c] * m + (1.f - m) * null_txt[c]);
    const float coord = lane < 4 ? boxes[(size_t)row * 4 + lane] : 0.f;
    for (int i = lane; i < nfreq * 4; i += 32) {
        const int k = i / 4, j = i % 4;
        const float x = __shfl_sync(0xffffffffu, coord, j);
        const float f = powf(100.0f, (float)k / (float)nfreq);
        const float a = f * x;
        const int base = k * 8;
        o[in_dim + base + j] = __float2half_rn(sinf(a) * m + (1.f - m) * null_pos[base + j]);
        o[in_dim + base + 4 + j] = __float2half_rn(cosf(a) * m + (1.f - m) * null_pos[base + 4 + j]);
    }
}
int posnet_input_launch(const float* boxes, const float* masks, const float* emb, const float* null_txt,
                        const float* null_pos, int rows, int in_dim, int nfreq, __half* out, cudaStream_t st) {
    if (nfreq * 4 > 32 * 4) {
        set_error("posnet: too many frequencies");
        return -1;
    }
    LTT_CUDA_OK(launch_k(posnet_input_kernel, dim3((rows + 3) / 4), dim3(128), 0, st, boxes, masks, emb, null_txt, null_pos, rows, in_dim, nfreq, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ relation fusion
// rects[b][i] = {top, bottom, left, right, valid}: reference truncation and first-break rules (attention.py:321-346)
__global__ void rela_rects_kernel(const float* __restrict__ boxes, const float* __restrict__ masks, int B, int mo, int h,
                                  int w, int* __restrict__ rects) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float nv = 0.f;
    for (int i = 0; i < mo; ++i) nv += masks[(size_t)b * mo + i];
    bool alive = true;
    for (int i = 0; i < mo; ++i) {
        const float* bx = boxes + ((size_t)b * mo + i) * 4;
        int l = (int)(bx[0] * (float)w), t = (int)(bx[1] * (float)h);
        int r = (int)fminf(bx[2] * (float)w, (float)w), bt = (int)fminf(bx[3] * (float)h, (float)h);
        const bool ok = alive && ((float)i < nv) && (l != r) && (t != bt);
        if (!ok) alive = false;
        l = max(0, min(l, w)); r = max(0, min(r, w)); t = max(0, min(t, h)); bt = max(0, min(bt, h));
        int* o = rects + ((size_t)b * mo + i) * 5;
        o[0] = t; o[1] = bt; o[2] = l; o[3] = r; o[4] = (ok && r > l && bt > t) ? 1 : 0;
    }
}
int rela_rects_launch(const float* boxes, const float* masks, int B, int mo, int h, int w, int* rects, cudaStream_t st) {
    LTT_CUDA_OK(launch_k(rela_rects_kernel, dim3((B + 31) / 32), dim3(32), 0, st, boxes, masks, B, mo, h, w, rects));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// (mean, rstd) of a token row for the kernels that normalise on the fly: either stored directly (slots == -1,
// layernorm_kernel's stats_out) or as the `slots` partial (sum, sum of squares) the producing GEMM's epilogue left
// (gemm_tc.cu, LayerNorm fold), summed in slot order.
__device__ __forceinline__ float2 row_mean_rstd(const RowStatSrc& s, size_t row) {
    const float2* p = s.p + row * (size_t)s.ld;
    if (s.slots < 0) return p[0];
    float sm = 0.f, sq = 0.f;
    for (int i = 0; i < s.slots; ++i) {
        const float2 t = p[i];
        sm += t.x;
        sq += t.y;
    }
    const float inv = 1.0f / (float)s.K;
    const float mean = sm * inv;
    return make_float2(mean, rsqrtf(fmaxf(sq * inv - mean * mean, 0.f) + s.eps));
}

// feats[b, i, :] = mean over the box of hid (fp32) -> fp16; zero for unused slots.  CTA = (64 channels, slot, b),
// RP_LANES pixel lanes x 16 channel quads (a box of a 64x64 latent holds up to 4096 pixels and only the valid slots do
// work, so the pixel loop is the critical path: 64 lanes x 4 pixels in flight), smem reduction over the pixel lanes.  hid is either a materialised
// fp32 tensor or (hid == nullptr) the LayerNorm of the fp16 tensor x16 evaluated on the fly from per-row statistics:
// hid[p][c] = (x16[p][c] - mean_p) * rstd_p * gamma[c] + beta[c]  (same expression / order as layernorm_kernel).
constexpr int RP_LANES = 64;
__global__ void __launch_bounds__(16 * RP_LANES) rela_pool_kernel(const float* __restrict__ hid, const __half* __restrict__ x16, const RowStatSrc stats,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, const int* __restrict__ rects,
                                 int mo, int w, int HW, int C, __half* __restrict__ feats) {
    pdl_launch_dependents();
    const int i = blockIdx.y, b = blockIdx.z, cq = threadIdx.x & 15, pl = threadIdx.x >> 4;
    const int c = blockIdx.x * 64 + cq * 4;
    float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!hid) {
        g4 = *reinterpret_cast<const float4*>(gamma + c);
        b4 = *reinterpret_cast<const float4*>(beta + c);
    }
    pdl_wait();
    const int* rc = rects + ((size_t)b * mo + i) * 5;
    __half* o = feats + ((size_t)b * mo + i) * C + c;
    if (!rc[4]) {
        if (pl == 0) *reinterpret_cast<uint2*>(o) = make_uint2(0, 0);
        return;
    }
    const int t = rc[0], bt = rc[1], l = rc[2], r = rc[3];
    const int rw = r - l, area = rw * (bt - t);
    float4 acc = make_float4(0, 0, 0, 0);
    auto value = [&](int p) -> float4 {
        const int y = t + p / rw, x = l + p % rw;
        const size_t row = (size_t)b * HW + y * w + x;
        if (hid) return *reinterpret_cast<const float4*>(hid + row * C + c);
        const uint2 u = *reinterpret_cast<const uint2*>(x16 + row * C + c);
        const float2 st = row_mean_rstd(stats, row);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4((f0.x - st.x) * st.y * g4.x + b4.x, (f0.y - st.x) * st.y * g4.y + b4.y,
                           (f1.x - st.x) * st.y * g4.z + b4.z, (f1.y - st.x) * st.y * g4.w + b4.w);
    };
    // four pixels in flight per thread; accumulation order = pixel order of this lane (as the single-pixel loop)
    for (int p = pl; p < area; p += 4 * RP_LANES) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (p + RP_LANES * k < area) v[k] = value(p + RP_LANES * k);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (p + RP_LANES * k < area) {
                acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
            }
    }
    __shared__ float4 red[16 * RP_LANES];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (pl < 16) {                                // two-level reduction over the pixel lanes, fixed order: deterministic
        for (int k = pl + 16; k < RP_LANES; k += 16) {
            const float4 v = red[k * 16 + cq];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        red[threadIdx.x] = acc;
    }
    __syncthreads();
    if (pl == 0) {
        for (int k = 1; k < 16; ++k) {
            const float4 v = red[k * 16 + cq];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        const float inv = 1.0f / (float)area;
        __half2 h0 = __floats2half2_rn(acc.x * inv, acc.y * inv), h1 = __floats2half2_rn(acc.z * inv, acc.w * inv);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(o) = u;
    }
}
int rela_pool_launch(const float* hid, const __half* x16, const RowStatSrc& stats, const float* gamma, const float* beta,
                     const int* rects, int B, int mo, int h, int w, int C, __half* feats, cudaStream_t st) {
    if (C % 64) {
        set_error("rela_pool: C %% 64 != 0 (C=%d)", C);
        return -1;
    }
    LTT_CUDA_OK(launch_k(rela_pool_kernel, dim3(dim3(C / 64, mo, B)), dim3(16 * RP_LANES), 0, st, hid, x16, stats, gamma, beta, rects, mo, w, h * w,
                         C, feats));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// out[b,p,:] = (hid + (1/mo) sum_{i: p in rect_i} feats[b,i,:] + x) / 2   (fp32).  feats == nullptr or b >= nb_feats
// -> no boxes (uncond half): out = (hid + x) / 2.  One CTA per pixel row; the row stays in registers and the
// LayerNorm that follows in the block (norm2, attention.py:437) is applied in the same kernel: ln16 = LN(out) in fp16
// (fp32 two-pass statistics, as layernorm_kernel).  C <= 8 * blockDim.x.
constexpr int RS_THREADS = 160;
__global__ void __launch_bounds__(RS_THREADS) rela_scatter_ln_kernel(
    const float* __restrict__ hid, const RowStatSrc stats, const float* __restrict__ gamma3,
    const float* __restrict__ beta3, const __half* __restrict__ x, const __half* __restrict__ feats,
    const int* __restrict__ rects, int nb_feats, int mo, int w, int HW, int C, float* __restrict__ out,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, __half* __restrict__ ln16) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x, b = row / HW, p = row % HW, y = p / w, xx = p % w;
    __shared__ int hit[32];
    __shared__ int nhit;
    __shared__ float red[2][RS_THREADS / 32];
    if (threadIdx.x < 32) {
        const int i = threadIdx.x;
        bool in = false;
        if (feats && b < nb_feats && i < mo) {
            const int* rc = rects + ((size_t)b * mo + i) * 5;
            in = rc[4] && y >= rc[0] && y < rc[1] && xx >= rc[2] && xx < rc[3];
        }
        const unsigned mask = __ballot_sync(0xffffffffu, in);
        if (in) hit[__popc(mask & ((1u << i) - 1))] = i;        // ascending slot order, as the sequential scan
        if (i == 0) nhit = __popc(mask);
    }
    __syncthreads();
    const float inv = 1.0f / (float)mo;
    const int c = threadIdx.x * 8;
    const bool act = c < C;
    float o[8];
    float s = 0.f;
    if (act) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < nhit; ++k) {
            const uint4 u = *reinterpret_cast<const uint4*>(feats + ((size_t)b * mo + hit[k]) * C + c);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                acc[2 * j] += f.x;
                acc[2 * j + 1] += f.y;
            }
        }
        const uint4 xu = *reinterpret_cast<const uint4*>(x + (size_t)row * C + c);
        const __half2* xh = reinterpret_cast<const __half2*>(&xu);
        float hv[8];
        if (hid) {
            const float4 h0 = *reinterpret_cast<const float4*>(hid + (size_t)row * C + c);
            const float4 h1 = *reinterpret_cast<const float4*>(hid + (size_t)row * C + c + 4);
            hv[0] = h0.x; hv[1] = h0.y; hv[2] = h0.z; hv[3] = h0.w; hv[4] = h1.x; hv[5] = h1.y; hv[6] = h1.z; hv[7] = h1.w;
        } else {      // hid = LayerNorm(x) from the row statistics (same expression / order as layernorm_kernel)
            const float2 st = row_mean_rstd(stats, (size_t)row);
            const float4 ga = *reinterpret_cast<const float4*>(gamma3 + c), gb = *reinterpret_cast<const float4*>(gamma3 + c + 4);
            const float4 ba = *reinterpret_cast<const float4*>(beta3 + c), bb = *reinterpret_cast<const float4*>(beta3 + c + 4);
            const float g3[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float b3[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xv = __half22float2(xh[j]);
                hv[2 * j] = (xv.x - st.x) * st.y * g3[2 * j] + b3[2 * j];
                hv[2 * j + 1] = (xv.y - st.x) * st.y * g3[2 * j + 1] + b3[2 * j + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 xv = __half22float2(xh[j]);
            o[2 * j] = ((hv[2 * j] + acc[2 * j] * inv) + xv.x) * 0.5f;
            o[2 * j + 1] = ((hv[2 * j + 1] + acc[2 * j + 1] * inv) + xv.y) * 0.5f;
        }
        float4* op = reinterpret_cast<float4*>(out + (size_t)row * C + c);
        op[0] = make_float4(o[0], o[1], o[2], o[3]);
        op[1] = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += o[j];
    }
    if (!ln16) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 16; k; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < RS_THREADS / 32; ++k) tot += red[0][k];
    const float mean = tot / C;
    float ss = 0.f;
    if (act) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = o[j] - mean;
            ss += d * d;
        }
    }
#pragma unroll
    for (int k = 16; k; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
    if (lane == 0) red[1][warp] = ss;
    __syncthreads();
    float tss = 0.f;
#pragma unroll
    for (int k = 0; k < RS_THREADS / 32; ++k) tss += red[1][k];
    const float rstd = rsqrtf(tss / C + eps);
    if (act) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            h[j] = __floats2half2_rn((o[2 * j] - mean) * rstd * g[2 * j] + bt[2 * j],
                                     (o[2 * j + 1] - mean) * rstd * g[2 * j + 1] + bt[2 * j + 1]);
        *reinterpret_cast<uint4*>(ln16 + (size_t)row * C + c) = *reinterpret_cast<uint4*>(h);
    }
}
// gamma == nullptr: scatter only (ln16 ignored).
int rela_scatter_launch(const float* hid, const RowStatSrc& stats, const float* gamma3, const float* beta3, const __half* x,
                        const __half* feats, const int* rects, int nb_feats, int B, int mo, int h, int w, int C, float* out,
                        const float* gamma, const float* beta, float eps, __half* ln16, cudaStream_t st) {
    if (mo > 32 || C % 8 || C > 8 * RS_THREADS) {
        set_error("rela_scatter: unsupported mo=%d C=%d", mo, C);
        return -1;
    }
    LTT_CUDA_OK(launch_k(rela_scatter_ln_kernel, dim3(B * h * w), dim3(RS_THREADS), 0, st, hid, stats, gamma3, beta3, x, feats, rects, nb_feats, mo, w,
                         h * w, C, out, gamma, beta, eps, gamma ? ln16 : (__half*)nullptr));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- relation cross-attention with the projections folded into the (step-invariant) relation keys / values.
// RelationCrossAttention (attention.py:315-359) attends 30 pooled box features to <= 32 relation tokens whose K / V are
// fixed for a whole sampling run.  With A[h,j,:] = scale * sum_{c in head h} Wq[c,:] k_j[c] and
// Bm[h,j,:] = sum_{c in head h} Wo[:,c] v_j[c] (rela_fold_kernel, once per conditioning) the chain
// to_q -> QK^T -> softmax -> PV -> to_out collapses to  logits = LN1(f) . A^T,  out = softmax(logits) . Bm + bias:
// no C x C weight is streamed per step and the branch's q-GEMM, attention and out-GEMM launches (three ~7 us floors on
// 30 rows) become part of one row kernel together with both LayerNorms and the gated residual.
// A, Bm: [G][heads*nrel][C] fp16.
// One block = 128 output channels of one (batch element, head): every weight element is read once per batch element
// and reused for all relations (accumulators in registers), K / V head slices sit in shared memory.
constexpr int RF_MAXREL = 32;
__global__ void __launch_bounds__(128) rela_fold_kernel(const __half* __restrict__ wq, const __half* __restrict__ wo,
                                                        const __half* __restrict__ kv, int nrel, int heads, int d, float scale,
                                                        __half* __restrict__ A, __half* __restrict__ Bm) {
    extern __shared__ float rf_smem[];
    float* ks = rf_smem;                 // [nrel][d]
    float* vs = ks + nrel * d;           // [nrel][d]
    const int C = heads * d, HJ = heads * nrel;
    const int h = blockIdx.y, g = blockIdx.z;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < nrel * d; i += blockDim.x) {
        const int j = i / d, cc = i - j * d;
        const __half* kr = kv + ((size_t)g * nrel + j) * 2 * C + h * d + cc;
        ks[i] = __half2float(kr[0]);
        vs[i] = __half2float(kr[C]);
    }
    __syncthreads();
    if (c >= C) return;
    float acc[RF_MAXREL];
#pragma unroll
    for (int j = 0; j < RF_MAXREL; ++j) acc[j] = 0.f;
    for (int cc0 = 0; cc0 < d; cc0 += 8) {     // A: Wq rows of head h, coalesced over the output channel c; d % 8 == 0
        __half wv[8];                          // eight independent loads in flight (the loop is L2-latency bound)
#pragma unroll
        for (int e = 0; e < 8; ++e) wv[e] = wq[(size_t)(h * d + cc0 + e) * C + c];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float w = __half2float(wv[e]);
#pragma unroll
            for (int j = 0; j < RF_MAXREL; ++j)
                if (j < nrel) acc[j] += w * ks[j * d + cc0 + e];
        }
    }
#pragma unroll
    for (int j = 0; j < RF_MAXREL; ++j)
        if (j < nrel) A[((size_t)g * HJ + h * nrel + j) * C + c] = __float2half_rn(acc[j] * scale);
#pragma unroll
    for (int j = 0; j < RF_MAXREL; ++j) acc[j] = 0.f;
    const __half* wr = wo + (size_t)c * C + h * d;        // Bm: row c of Wo, columns of head h (16-byte vectors)
    for (int c8 = 0; c8 < (d >> 3); c8 += 4) {
        uint4 u4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (c8 + q < (d >> 3)) u4[q] = *reinterpret_cast<const uint4*>(wr + (c8 + q) * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c8 + q >= (d >> 3)) continue;
            const __half2* h2 = reinterpret_cast<const __half2*>(&u4[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                const int cc = (c8 + q) * 8 + 2 * e;
#pragma unroll
                for (int j = 0; j < RF_MAXREL; ++j)
                    if (j < nrel) {
                        acc[j] += f.x * vs[j * d + cc];
                        acc[j] += f.y * vs[j * d + cc + 1];
                    }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RF_MAXREL; ++j)
        if (j < nrel) Bm[((size_t)g * HJ + h * nrel + j) * C + c] = __float2half_rn(acc[j]);
}
int rela_fold_launch(const __half* wq, const __half* wo, const __half* kv, int G, int nrel, int heads, int d, float scale,
                     __half* A, __half* Bm, cudaStream_t st) {
    const int C = heads * d;
    const size_t smem = (size_t)2 * nrel * d * sizeof(float);
    if (nrel > RF_MAXREL || d % 8 || smem > 48 * 1024) {
        set_error("rela_fold: unsupported nrel=%d d=%d", nrel, d);
        return -1;
    }
    rela_fold_kernel<<<dim3((C + 127) / 128, heads, G), 128, smem, st>>>(wq, wo, kv, nrel, heads, d, scale, A, Bm);
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// One CTA per (pooled feature row, head):  ln = fp16(LN1(f)) (recomputed per head: the row is 2.5 KB);
// logits[j] = fp16(ln . A[g,h,j]);  p = fp16(softmax_j);  partial[h] = sum_j p_j Bm[g,h,j] -> fp32 scratch.  The last
// CTA of a row to finish (ticket counter) adds the heads' partials in head order (deterministic), applies
// y = fp16(sum + bias), f2 = fp16(f + fp16(gate * y)), ln2 = fp16(LN2(f2)) and re-arms the counter.  Rounding points
// follow the fp16 autocast reference except that q and the per-head attention output are never materialised.
// rows x heads CTAs keep the L2 round trips per CTA at three instead of fifteen for a one-CTA-per-row layout.
constexpr int RA_THREADS = 256;
__global__ void __launch_bounds__(RA_THREADS) rela_attn_fused_kernel(
    const __half* __restrict__ feats, int rows_per_g, int C, int heads, int nrel, const __half* __restrict__ A,
    const __half* __restrict__ Bm, const float* __restrict__ bias, float gate, const float* __restrict__ g1,
    const float* __restrict__ b1, const float* __restrict__ g2, const float* __restrict__ b2, float eps,
    float* __restrict__ scratch, int* __restrict__ tickets, __half* __restrict__ feats2, __half* __restrict__ ln2out) {
    pdl_launch_dependents();
    extern __shared__ float ra_smem[];
    float* xs = ra_smem;                 // [C] feature row, later f2
    float* ls = xs + C;                  // [max(C, 2048)] LN1 output (fp16-rounded), later partial sums of the out phase
    float* lg = ls + max(C, 8 * RA_THREADS);   // [nrel] logits, then probabilities
    __shared__ float red[RA_THREADS / 32];
    __shared__ int is_last;
    const int row = blockIdx.x, h = blockIdx.y, g = row / rows_per_g;
    const int HJ = heads * nrel, nvec = C >> 3;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto block_sum = [&](float v) -> float {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();                 // protects `red` from the previous use
        if (lane == 0) red[warp] = v;
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < RA_THREADS / 32; ++k) t += red[k];
        return t;
    };
    // ---- step-invariant operands first: the folded relation keys / values of this (sample, head) and the norm parameters do
    // not depend on the predecessor grid, so their L2 round trips overlap its tail and the LayerNorm below instead of sitting
    // on the kernel's dependent chain (load row -> LN1 -> logits -> softmax -> p . Bm -> ticket -> LN2)
    const __half* Ah = A + ((size_t)g * HJ + (size_t)h * nrel) * C;
    const __half* Bh = Bm + ((size_t)g * HJ + (size_t)h * nrel) * C;
    const int P = max(1, RA_THREADS / nvec);
    uint4 ua[5], ub[8];
    if (warp < nrel) {                   // logits: warp j handles relation j (nrel <= 8 warps; a second round reloads below)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int cv = lane + 32 * i;
            if (cv < nvec) ua[i] = *reinterpret_cast<const uint4*>(Ah + (size_t)warp * C + cv * 8);
        }
    }
    {
        const int cv = tid % nvec, part = tid / nvec;
        if (part < P) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = part + i * P;
                if (j < nrel) ub[i] = *reinterpret_cast<const uint4*>(Bh + (size_t)j * C + cv * 8);
            }
        }
    }
    float g1r[5], b1r[5];                // LN1 parameters of this thread's channels (C <= 5 * RA_THREADS)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int c = tid + i * RA_THREADS;
        if (c < C) {
            g1r[i] = g1[c];
            b1r[i] = b1[c];
        }
    }
    pdl_wait();
    // ---- load + LN1
    float s = 0.f;
    for (int c = tid; c < C; c += RA_THREADS) {
        const float v = __half2float(feats[(size_t)row * C + c]);
        xs[c] = v;
        s += v;
    }
    const float mean = block_sum(s) / C;
    float ss = 0.f;
    for (int c = tid; c < C; c += RA_THREADS) {
        const float dlt = xs[c] - mean;
        ss += dlt * dlt;
    }
    const float rstd = rsqrtf(block_sum(ss) / C + eps);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int c = tid + i * RA_THREADS;
        if (c < C) ls[c] = r16f((xs[c] - mean) * rstd * g1r[i] + b1r[i]);
    }
    __syncthreads();
    // ---- logits of this head: one warp per relation, lanes over C in 16-byte vectors (<= 5 loads in flight per lane)
    for (int j = warp; j < nrel; j += RA_THREADS / 32) {
        uint4 u[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int cv = lane + 32 * i;
            if (cv < nvec) u[i] = j == warp ? ua[i] : *reinterpret_cast<const uint4*>(Ah + (size_t)j * C + cv * 8);
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int cv = lane + 32 * i;
            if (cv < nvec) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(h2[k]);
                    acc += ls[cv * 8 + 2 * k] * f.x + ls[cv * 8 + 2 * k + 1] * f.y;
                }
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) lg[j] = r16f(acc);
    }
    __syncthreads();
    if (tid == 0) {                      // softmax over the relations
        float mx = -INFINITY;
        for (int j = 0; j < nrel; ++j) mx = fmaxf(mx, lg[j]);
        float sum = 0.f;
        for (int j = 0; j < nrel; ++j) {
            const float e = __expf(lg[j] - mx);
            lg[j] = e;
            sum += e;
        }
        for (int j = 0; j < nrel; ++j) lg[j] = r16f(lg[j] / sum);
    }
    __syncthreads();
    // ---- this head's share of p . Bm: relations split over P thread groups, all loads of a thread in flight together
    // (the first round was fetched ahead of the wait)
    {
        const int cv = tid % nvec, part = tid / nvec;
        if (part < P) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int jb = part; jb < nrel; jb += P * 8) {
                uint4 u[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = jb + i * P;
                    if (j < nrel) u[i] = jb == part ? ub[i] : *reinterpret_cast<const uint4*>(Bh + (size_t)j * C + cv * 8);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = jb + i * P;
                    if (j < nrel) {
                        const float p = lg[j];
                        const __half2* h2 = reinterpret_cast<const __half2*>(&u[i]);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 f = __half22float2(h2[k]);
                            acc[2 * k] += p * f.x;
                            acc[2 * k + 1] += p * f.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) ls[part * C + cv * 8 + k] = acc[k];      // partials [P][C]; LN1 output is dead now
        }
    }
    __syncthreads();
    float* mine = scratch + ((size_t)row * heads + h) * C;
    for (int c = tid; c < C; c += RA_THREADS) {
        float acc = 0.f;
        for (int pp = 0; pp < P; ++pp) acc += ls[pp * C + c];
        mine[c] = acc;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(&tickets[row], 1) == heads - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- last CTA of the row: heads in fixed order, bias, gated residual, LN2
    const float* all = scratch + (size_t)row * heads * C;
    float s2 = 0.f;
    for (int c = tid; c < C; c += RA_THREADS) {
        float acc = 0.f;
        for (int hh = 0; hh < heads; ++hh) acc += __ldcg(all + (size_t)hh * C + c);
        const float y = r16f(acc + bias[c]);
        const float f2 = r16f(xs[c] + r16f(gate * y));
        xs[c] = f2;
        s2 += f2;
    }
    const float mean2 = block_sum(s2) / C;
    float ss2 = 0.f;
    for (int c = tid; c < C; c += RA_THREADS) {
        const float dlt = xs[c] - mean2;
        ss2 += dlt * dlt;
    }
    const float rstd2 = rsqrtf(block_sum(ss2) / C + eps);
    for (int cv = tid; cv < nvec; cv += RA_THREADS) {
        __half2 o1[4], o2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = cv * 8 + 2 * k;
            o1[k] = __floats2half2_rn(xs[c], xs[c + 1]);
            o2[k] = __floats2half2_rn((xs[c] - mean2) * rstd2 * g2[c] + b2[c], (xs[c + 1] - mean2) * rstd2 * g2[c + 1] + b2[c + 1]);
        }
        *reinterpret_cast<uint4*>(feats2 + (size_t)row * C + cv * 8) = *reinterpret_cast<uint4*>(o1);
        *reinterpret_cast<uint4*>(ln2out + (size_t)row * C + cv * 8) = *reinterpret_cast<uint4*>(o2);
    }
    if (tid == 0) tickets[row] = 0;      // re-armed for the next launch (graph replays)
}
// scratch: >= G*rows_per_g*heads*C floats; tickets: >= G*rows_per_g ints, zero before the first launch.
int rela_attn_fused_launch(const __half* feats, int G, int rows_per_g, int C, int heads, int nrel, const __half* A,
                           const __half* Bm, const float* bias, float gate, const float* g1, const float* b1,
                           const float* g2, const float* b2, float eps, float* scratch, int* tickets, __half* feats2,
                           __half* ln2out, cudaStream_t st) {
    const size_t smem = ((size_t)C + std::max<size_t>(C, 8 * RA_THREADS) + (size_t)nrel) * sizeof(float);
    if (C % 8 || C > 1280 || nrel > 32 || smem > 48 * 1024) {
        set_error("rela_attn_fused: unsupported C=%d heads=%d nrel=%d", C, heads, nrel);
        return -1;
    }
    LTT_CUDA_OK(launch_k(rela_attn_fused_kernel, dim3(G * rows_per_g, heads), dim3(RA_THREADS), smem, st, feats, rows_per_g, C, heads, nrel,
                         A, Bm, bias, gate, g1, b1, g2, b2, eps, scratch, tickets, feats2, ln2out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// 30-query x (<= 32)-key cross attention of the relation path.  One block per (batch element, head): q / k / v head
// slices are staged in shared memory with 16-byte loads (the previous one-warp-per-query version walked them with
// dependent 4-byte global loads and was pure latency), then one warp per query: lane j owns key j for the logits,
// softmax by warp shuffles, lanes own channels for P V.  Rounding points follow the fp16 autocast reference
// (attention.py:127-141): logits and probabilities rounded to fp16.
// q: [B, nq, ldq], k/v: [B, nk, ldkv] fp16 row-major, out [B, nq, heads*d] fp16.
constexpr int SA_THREADS = 256;
__global__ void __launch_bounds__(SA_THREADS) small_attn_kernel(const __half* __restrict__ q, int ldq, const __half* __restrict__ k,
                                                               const __half* __restrict__ v, int ldkv, int nq, int nk, int heads,
                                                               int d, float scale, __half* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sa_smem[];
    const int hh = blockIdx.x, b = blockIdx.y;
    const int ds = d + 4;                                   // padded row stride (keeps 16-byte alignment, spreads banks)
    float* qs = sa_smem;                                    // [nq][ds]
    float* ks = qs + nq * ds;                               // [nk][ds]
    float* vs = ks + nk * ds;                               // [nk][ds]
    float* ps = vs + nk * ds;                               // [nq][32] logits, then probabilities
    const int dv = d >> 3;                                  // 16-byte vectors per head row (host: d % 8 == 0)
    for (int i = threadIdx.x; i < (nq + 2 * nk) * dv; i += blockDim.x) {
        const int row = i / dv, cv = i - row * dv;
        const __half* src;
        float* dst;
        if (row < nq) { src = q + ((size_t)b * nq + row) * ldq + hh * d; dst = qs + row * ds; }
        else if (row < nq + nk) { src = k + ((size_t)b * nk + (row - nq)) * ldkv + hh * d; dst = ks + (row - nq) * ds; }
        else { src = v + ((size_t)b * nk + (row - nq - nk)) * ldkv + hh * d; dst = vs + (row - nq - nk) * ds; }
        const uint4 u = *reinterpret_cast<const uint4*>(src + cv * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
        float4* d4 = reinterpret_cast<float4*>(dst + cv * 8);
        d4[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
        d4[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    __syncthreads();
    // logits: one (query, key) pair per thread; pairwise accumulation order as the reference-checked first version
    for (int idx = threadIdx.x; idx < nq * nk; idx += blockDim.x) {
        const int qi = idx / nk, j = idx - qi * nk;
        const float4* qr = reinterpret_cast<const float4*>(qs + qi * ds);
        const float4* kr = reinterpret_cast<const float4*>(ks + j * ds);
        float acc = 0.f;
        for (int c = 0; c < (d >> 2); ++c) {
            const float4 a = qr[c], bb = kr[c];
            acc += a.x * bb.x + a.y * bb.y;
            acc += a.z * bb.z + a.w * bb.w;
        }
        ps[qi * 32 + j] = r16f(r16f(acc) * scale);   // reference: fp16 einsum result times scale in fp16
    }
    __syncthreads();
    if (threadIdx.x < nq) {                                 // softmax of one query row (nk <= 32 keys)
        float* pr = ps + threadIdx.x * 32;
        float mx = -INFINITY;
        for (int j = 0; j < nk; ++j) mx = fmaxf(mx, pr[j]);
        float sum = 0.f;
        for (int j = 0; j < nk; ++j) {
            const float e = __expf(pr[j] - mx);
            pr[j] = e;
            sum += e;
        }
        for (int j = 0; j < nk; ++j) pr[j] = r16f(pr[j] / sum);
    }
    __syncthreads();
    const int C = heads * d;
    for (int idx = threadIdx.x; idx < nq * d; idx += blockDim.x) {
        const int qi = idx / d, c = idx - qi * d;
        float acc = 0.f;
        for (int j = 0; j < nk; ++j) acc += ps[qi * 32 + j] * vs[j * ds + c];
        out[((size_t)b * nq + qi) * C + hh * d + c] = __float2half_rn(acc);
    }
}
int small_attn_launch(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int nq, int nk,
                      int heads, int d, float scale, __half* out, cudaStream_t st) {
    const size_t smem = ((size_t)(nq + 2 * nk) * (d + 4) + (size_t)nq * 32) * sizeof(float);
    if (nk > 32 || (d & 7) || ldq % 8 || ldkv % 8 || smem > 48 * 1024) {
        set_error("small_attn: nq=%d nk=%d d=%d unsupported", nq, nk, d);
        return -1;
    }
    LTT_CUDA_OK(launch_k(small_attn_kernel, dim3(heads, B), dim3(SA_THREADS), smem, st, q, ldq, k, v, ldkv, nq, nk, heads, d, scale, out));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ misc
__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __float2half_rn(in[i]);
}
int cast_f32_f16_launch(const float* in, __half* out, size_t n, cudaStream_t st) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(cast_f32_f16_kernel, dim3(blocks), dim3(256), 0, st, in, out, n));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void cast_f16_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = __half2float(in[i]);
}
int cast_f16_f32_launch(const __half* in, float* out, size_t n, cudaStream_t st) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(cast_f16_f32_kernel, dim3(blocks), dim3(256), 0, st, in, out, n));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// generic strided 2-D fp16 copy: dst[b][r][c] = src[b][r][c], rows x cols per batch element
__global__ void copy2d_kernel(const __half* __restrict__ src, size_t sb, int sld, __half* __restrict__ dst, size_t db,
                              int dld, int B, int rows, int cols) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)B * rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        const size_t p = i / cols;
        const int r = (int)(p % rows), b = (int)(p / rows);
        dst[b * db + (size_t)r * dld + c] = src[b * sb + (size_t)r * sld + c];
    }
}
int copy2d_launch(const __half* src, size_t sb, int sld, __half* dst, size_t db, int dld, int B, int rows, int cols,
                  cudaStream_t st) {
    const size_t total = (size_t)B * rows * cols;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
    LTT_CUDA_OK(launch_k(copy2d_kernel, dim3(blocks), dim3(256), 0, st, src, sb, sld, dst, db, dld, B, rows, cols));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// CFG + PLMS update (plms.py:110-163), one launch per sampler step.  All eps tensors hold fp16-representable values
// (the autocast reference returns half), arithmetic on them is rounded to fp16 per op exactly as eager PyTorch does;
// x stays fp32.  mode: 0 = first half of step 0 (Euler predictor: writes x_pred for the second evaluation),
// 1 = second half of step 0 (e' = (e_t + e_next)/2), 2..4 = AB2/AB3/AB4.
__global__ void plms_update_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u, float guidance,
                                   int use_cfg, int mode, const float* __restrict__ x, float* __restrict__ e_t_out,
                                   const float* __restrict__ e_first, const float* __restrict__ old1,
                                   const float* __restrict__ old2, const float* __restrict__ old3, float a_t,
                                   float a_prev, float sqrt_1m_at, float* __restrict__ x_out, size_t n) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float e = eps_c[i];
        if (use_cfg) {
            const float u = eps_u[i];
            e = r16f(u + r16f(guidance * r16f(e - u)));
        }
        float ep;
        if (mode == 0) {
            e_t_out[i] = e;
            ep = e;
        } else if (mode == 1) {
            ep = r16f(r16f(e_first[i] + e) * 0.5f);     // e_t kept from mode 0 in e_first; old_eps gets e_first
        } else {
            e_t_out[i] = e;
            if (mode == 2) ep = r16f(r16f(r16f(3.f * e) - old1[i]) * 0.5f);
            else if (mode == 3) ep = r16f(r16f(r16f(r16f(23.f * e) - r16f(16.f * old1[i])) + r16f(5.f * old2[i])) / 12.f);
            else ep = r16f(r16f(r16f(r16f(r16f(55.f * e) - r16f(59.f * old1[i])) + r16f(37.f * old2[i])) - r16f(9.f * old3[i])) / 24.f);
        }
        const float pred_x0 = (x[i] - sqrt_1m_at * ep) / sqrtf(a_t);
        const float dir = sqrtf(1.0f - a_prev) * ep;
        x_out[i] = sqrtf(a_prev) * pred_x0 + dir;
    }
}
int plms_update_launch(const float* eps_c, const float* eps_u, float guidance, int use_cfg, int mode, const float* x,
                       float* e_t_out, const float* e_first, const float* old1, const float* old2, const float* old3,
                       float a_t, float a_prev, float sqrt_1m_at, float* x_out, size_t n, cudaStream_t st) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 4);
    LTT_CUDA_OK(launch_k(plms_update_kernel, dim3(blocks), dim3(256), 0, st, eps_c, eps_u, guidance, use_cfg, mode, x, e_t_out, e_first, old1, old2,
                                                old3, a_t, a_prev, sqrt_1m_at, x_out, n));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// UNet inputs of one sampler evaluation: x_in = [x ; x] (cond and uncond rows see the same latent, plms.py:116-122) and
// the timestep vector, written on the device so the sampler loop never synchronises with the host.
// ev_src != nullptr: the time-embedding row of this evaluation (precomputed for every sampler step, see
// ltt_plms_sample) is broadcast to the B rows of the per-evaluation buffer the ResBlock epilogues read.
__global__ void plms_prep_kernel(const float* __restrict__ x, float* __restrict__ x_in, size_t n, int copies,
                                 float* __restrict__ t_in, int B, float tval, const uint4* __restrict__ ev_src,
                                 uint4* __restrict__ ev_dst, int ev_vecs) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (size_t)B) t_in[gid] = tval;
    for (size_t i = gid; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        for (int c = 0; c < copies; ++c) x_in[(size_t)c * n + i] = v;
    }
    if (ev_src)
        for (size_t i = gid; i < (size_t)ev_vecs; i += (size_t)gridDim.x * blockDim.x) {
            const uint4 v = ev_src[i];
            for (int b = 0; b < B; ++b) ev_dst[(size_t)b * ev_vecs + i] = v;
        }
}
int plms_prep_launch(const float* x, float* x_in, size_t n, int copies, float* t_in, int B, float tval, const __half* ev_src,
                     __half* ev_dst, int ev_len, cudaStream_t st) {
    if (ev_src && ev_len % 8) {
        set_error("plms_prep: embedding row length %d is not a multiple of 8", ev_len);
        return -1;
    }
    const int blocks = (int)std::min<size_t>((std::max<size_t>(n, (size_t)B) + 255) / 256, 148 * 4);
    LTT_CUDA_OK(launch_k(plms_prep_kernel, dim3(blocks), dim3(256), 0, st, x, x_in, n, copies, t_in, B, tval,
                         reinterpret_cast<const uint4*>(ev_src), reinterpret_cast<uint4*>(ev_dst), ev_len / 8));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// t_tab[i] = v[i]: the sampler's evaluation timesteps, passed by value (no host buffer to keep alive, no staging copy)
struct TVals { float v[64]; };
__global__ void fill_tvals_kernel(TVals tv, int n, float* __restrict__ out) {
    if ((int)threadIdx.x < n) out[threadIdx.x] = tv.v[threadIdx.x];
}
int fill_tvals_launch(const float* host_vals, int n, float* out, cudaStream_t st) {
    if (n > 64) {
        set_error("fill_tvals: more than 64 values");
        return -1;
    }
    TVals tv;
    for (int i = 0; i < 64; ++i) tv.v[i] = i < n ? host_vals[i] : 0.f;
    fill_tvals_kernel<<<1, 64, 0, st>>>(tv, n, out);
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ weight repacking
// conv weight [O, Cin, taps] fp32 (OIHW flattened) -> fp16 dst[o*Kdst + koff + tap*Cs + c] = w[o][cstart + c][tap]
__global__ void pack_conv_kernel(const float* __restrict__ w, int O, int Cin, int taps, int cstart, int Cs,
                                 __half* __restrict__ dst, int Kdst, int koff) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)O * taps * Cs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cs);
        const size_t p = i / Cs;
        const int t = (int)(p % taps), o = (int)(p / taps);
        dst[(size_t)o * Kdst + koff + t * Cs + c] = __float2half_rn(w[((size_t)o * Cin + cstart + c) * taps + t]);
    }
}
int pack_conv_launch(const float* w, int O, int Cin, int taps, int cstart, int Cs, __half* dst, int Kdst, int koff,
                     cudaStream_t st) {
    const size_t total = (size_t)O * taps * Cs;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(pack_conv_kernel, dim3(blocks), dim3(256), 0, st, w, O, Cin, taps, cstart, Cs, dst, Kdst, koff));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

// linear weight rows [rows, K] fp32 -> fp16 dst rows starting at row_off; geglu != 0 interleaves per 128-row tile:
// packed row p = tile*128 + h*64 + i  <-  source row h*(rows/2) + tile*64 + i   (h = 0 value, 1 gate)
__global__ void pack_rows_kernel(const float* __restrict__ w, int rows, int K, __half* __restrict__ dst, int row_off,
                                 int geglu, const float* __restrict__ colscale) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)rows * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int p = (int)(i / K);
        int src = p;
        if (geglu) {
            const int tile = p / 128, h = (p % 128) / 64, ii = p % 64;
            src = h * (rows / 2) + tile * 64 + ii;
        }
        const float v = w[(size_t)src * K + k];
        dst[((size_t)row_off + p) * K + k] = __float2half_rn(colscale ? v * colscale[k] : v);
    }
}
// LayerNorm fold (gemm_tc.cu): for W' = fp16(W * gamma) the consumer GEMM needs per output row n
//   s[n] = sum_k W'[n, k]   and   c[n] = bias[n] + sum_k beta[k] * fp16(W[n, k]).   One warp per row (original row order).
__global__ void ln_fold_vectors_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ bias, int N, int K, float* __restrict__ s_out,
                                       float* __restrict__ c_out) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.f, c = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float v = w[(size_t)n * K + k];
        s += r16f(v * gamma[k]);
        c = fmaf(beta[k], r16f(v), c);
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
        s_out[n] = s;
        c_out[n] = c + (bias ? bias[n] : 0.f);
    }
}
int ln_fold_vectors_launch(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K, float* s_out,
                           float* c_out, cudaStream_t st) {
    ln_fold_vectors_kernel<<<(N + 7) / 8, 256, 0, st>>>(w, gamma, beta, bias, N, K, s_out, c_out);
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}
int pack_rows_launch(const float* w, int rows, int K, __half* dst, int row_off, int geglu, cudaStream_t st, const float* colscale) {
    if (geglu && rows % 128) {
        set_error("pack_rows: GEGLU rows %% 128 != 0 (%d)", rows);
        return -1;
    }
    const size_t total = (size_t)rows * K;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
    LTT_CUDA_OK(launch_k(pack_rows_kernel, dim3(blocks), dim3(256), 0, st, w, rows, K, dst, row_off, geglu, colscale));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ltt
