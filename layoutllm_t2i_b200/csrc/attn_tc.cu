// Flash-style attention core on tcgen05 for sm_100a: S = Q K^T and O += P V both run on the tensor cores with the
// S tiles (double buffered) and the O accumulator resident in TMEM; softmax is online in the exp2 domain with a
// lazy (thresholded) rescale of O, so O is touched by the CUDA cores only when a row maximum jumps.
//
// One CTA = 128 query rows of one (batch, head).  warp 0: TMA producer (Q once, K / V^T tiles double buffered),
// warp 1: tcgen05.mma issuer (both warp-converged with elected issue), then NWG softmax warpgroups: NWG threads per
// query row (TMEM lane), each owning BKV / NWG columns of the score tile.  P is written to shared memory in the
// 128-B-swizzled K-major layout and consumed as the A operand of the PV MMA; V arrives transposed ([C, keys], keys
// contiguous -- written that way by the QKV projection's epilogue) so it is a K-major B operand.
// Head dims 40 / 80 / 160 run as K = 48 / 80 / 160 for QK^T (zero padded columns) and N = 48 / 80 / 160 for PV.
//
// Replaces the einsum / softmax / einsum triple of the reference's attention modules
// (/root/reference/GLIGEN/ldm/modules/attention.py:127-141 CrossAttention, :164-176 SelfAttention), which
// materialise the [B*8, N, N] score matrix.
#include <cstdio>
#include <cstdlib>

#include "ltt_kernels.h"
#include "ltt_ptx.cuh"

namespace ltt {

constexpr int att_threads(int nwg) { return 64 + 128 * nwg; }    // TMA warp, MMA warp, NWG softmax warpgroups

struct AttnDeviceArgs {
    CUtensorMap qmap, kmap, vmap;
    int nq, nk, dhead, dpad;
    __half* out;
    int ldo;
    float scale_log2;
};

// PTM: the probabilities P live in tensor memory (tcgen05.st by the softmax threads, A-from-TMEM operand of the PV MMA)
// instead of a swizzled shared-memory tile: per 128 x 128 score tile that takes 64 KB of shared-memory traffic (32 KB of
// st.shared + 32 KB of MMA operand reads, half of the tile's total) off the 128 B/clk shared-memory port.
template <int DPAD, int DV, int BKV, int SBUF, int PBUF, bool PTM = false>
struct AttnCfg {
    static constexpr int NKC = DPAD / 64;           // 64-column chunks of Q / K rows
    static constexpr int KSTEPS = DV / 16;          // k16 steps of QK^T (d rounded up to 16)
    static constexpr int NVC = BKV / 64;            // 64-key chunks of a KV tile
    static constexpr int Q_BYTES = NKC * 128 * 128;
    static constexpr int K_BYTES = NKC * BKV * 128;
    static constexpr int V_BYTES = NVC * DV * 128;
    static constexpr int P_BYTES = NVC * 128 * 128;
    static constexpr int OFF_K = Q_BYTES;
    static constexpr int OFF_V = OFF_K + 2 * K_BYTES;
    static constexpr int OFF_P = OFF_V + 2 * V_BYTES;
    static constexpr int OFF_BAR = OFF_P + (PTM ? 0 : PBUF * P_BYTES);
    static constexpr int OFF_XCH = OFF_BAR + 256;                 // [4][128] floats: row max / row sum exchange between warpgroups
    static constexpr int TOTAL = OFF_XCH + 2048 + 1024;
    static constexpr int TM_S = 0, TM_P = SBUF * BKV;                       // P: BKV / 2 packed columns per buffer
    static constexpr int TM_O = TM_P + (PTM ? PBUF * BKV / 2 : 0);
    static constexpr int TM_USED = TM_O + DV;
    static constexpr int TM_COLS = TM_USED <= 64 ? 64 : TM_USED <= 128 ? 128 : TM_USED <= 256 ? 256 : 512;
};

template <bool V>
struct BoolTag {
    static constexpr bool value = V;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA / ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial on [-0.5, 0.5], relative error
// 7.5e-5 -- a sixth of an fp16 ulp, P is rounded to fp16 anyway): the MUFU unit delivers 16 ex2 per clock per SM
// (measured: tools/ubench/tmem_mufu.cu, 15.3-15.5), i.e. 1024+ cycles for a 128 x 128 score tile against ~400 cycles of
// tensor-core work at d = 40, so the softmax is MUFU bound; evaluating every POLY-th exponential here instead moves that
// share of the load onto pipes that are otherwise ~35 % busy (the idea of FlashAttention-4's software exp2).
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);                        // masked (-inf) scores -> 2^-126 ~ 0
    const float t = x + 12582912.0f;              // 1.5 * 2^23: the integer nearest to x lands in the low mantissa bits
    const float n = t - 12582912.0f;
    const float f = x - n;
    float p = fmaf(0.055171624f, f, 0.24261113f);
    p = fmaf(p, f, 0.69326097f);
    p = fmaf(p, f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));      // * 2^n through the exponent field
}

// SBUF = number of S accumulators in TMEM: 2 lets QK^T of tile j+1 overlap the softmax of tile j inside one CTA;
// 1 (with MINB = 2 CTAs per SM, 256 TMEM columns each) gets the same overlap from the co-resident CTA and doubles
// the number of softmax warps per SM -- the d=40 level is bound by the softmax warps' issue rate, not by the tensor pipe.
// PBUF = number of P tiles in shared memory: with 2 (and SBUF = 2) the softmax of tile j+1 never waits for the tensor
// core -- S(j+1) was produced while tile j was exponentiated and P(j+1) goes to the other buffer while PV(j) runs.
// NWG = softmax warpgroups: with 2 the columns of a score tile are split between two threads per query row (half the
// serial TMEM-load -> exp -> store chain per tile); the pair agrees on the running reference through shared memory.
// POLY = k > 0: every k-th exponential of a row goes through ex2_poly instead of MUFU.EX2.
// CAUSAL: key j is visible to query i iff j <= i (compile-time, so that the UNet's kernels keep their uniform tile bounds).
template <int DPAD, int DV, int BKV, int SBUF, int PBUF, int MINB, int NWG, int POLY, bool PTM, bool CAUSAL = false>
__global__ void __launch_bounds__(att_threads(NWG), MINB) attn_tc_kernel(const __grid_constant__ AttnDeviceArgs args) {
    using C = AttnCfg<DPAD, DV, BKV, SBUF, PBUF, PTM>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // 2
    uint64_t* k_empty = bars + 3;       // 2
    uint64_t* v_full = bars + 5;        // 2
    uint64_t* v_empty = bars + 7;       // 2
    uint64_t* s_full = bars + 9;        // 2
    uint64_t* s_empty = bars + 11;      // 2  (count 128)
    uint64_t* p_full = bars + 13;       // 2  (count 128)
    uint64_t* pv_done = bars + 15;      // 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    // warp index through a shuffle: warp-uniform for ptxas, so the producer / MMA-issuer loops below stay on the uniform
    // datapath (their TMA / tcgen05 operations are issued by one elected lane inside the "_elect" wrappers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
    // causal: KV tiles entirely to the right of this CTA's last query row are never visited (all three roles agree on nt)
    const int nt = CAUSAL ? min((args.nk + BKV - 1) / BKV, (q0 + 127) / BKV + 1) : (args.nk + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.qmap);
        tma_prefetch_desc(&args.kmap);
        tma_prefetch_desc(&args.vmap);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 128 * NWG);
            mbar_init(&p_full[i], 128 * NWG);
            mbar_init(&pv_done[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<C::TM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_launch_dependents();     // programmatic dependent launch: see ltt_ptx.cuh
    pdl_wait();
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);     // barrier i lives at bar_base + 8 * i (same order as the pointers above)

    if (warp == 0) {
        mbar_expect_tx_elect(bar_base + 0, C::Q_BYTES);
        for (int c = 0; c < C::NKC; ++c)
            tma_load_3d_elect(smem_base + c * (128 * 128), &args.qmap, bar_base + 0, head * DPAD + c * 64, q0, b);
        for (int j = 0; j < nt; ++j) {
            const int st = j & 1;
            const uint32_t par = ((j >> 1) & 1) ^ 1;
            mbar_wait(&k_empty[st], par);
            mbar_expect_tx_elect(bar_base + 8 * (1 + st), C::K_BYTES);
            for (int c = 0; c < C::NKC; ++c)
                tma_load_3d_elect(smem_base + C::OFF_K + st * C::K_BYTES + c * (BKV * 128), &args.kmap, bar_base + 8 * (1 + st),
                                  head * DPAD + c * 64, j * BKV, b);
            mbar_wait(&v_empty[st], par);
            mbar_expect_tx_elect(bar_base + 8 * (5 + st), C::V_BYTES);
            for (int c = 0; c < C::NVC; ++c)
                tma_load_3d_elect(smem_base + C::OFF_V + st * C::V_BYTES + c * (DV * 128), &args.vmap, bar_base + 8 * (5 + st),
                                  j * BKV + c * 64, head * args.dhead, b);
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_qk = umma_idesc_f16(BKV);
        constexpr uint32_t idesc_pv = umma_idesc_f16(DV);
        const uint32_t sq = smem_base;
        const uint32_t sk = smem_base + C::OFF_K;
        const uint32_t sv = smem_base + C::OFF_V;
        const uint32_t sp = smem_base + C::OFF_P;
        auto issue_qk = [&](int j) {
            const int st = j & 1, sb = j % SBUF;
            mbar_wait(&k_full[st], (j >> 1) & 1);
            if (j >= SBUF) mbar_wait(&s_empty[sb], ((j / SBUF) - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < C::KSTEPS; ++kk) {
                const uint64_t ad = umma_desc_sw128(sq + (kk >> 2) * (128 * 128)) + 2 * (kk & 3);
                const uint64_t bd = umma_desc_sw128(sk + st * C::K_BYTES + (kk >> 2) * (BKV * 128)) + 2 * (kk & 3);
                umma_f16_elect(tmem_base + C::TM_S + sb * BKV, ad, bd, idesc_qk, kk != 0);
            }
            umma_commit_elect(bar_base + 8 * (3 + st));      // k_empty[st]
            umma_commit_elect(bar_base + 8 * (9 + sb));      // s_full[sb]
        };
        mbar_wait(q_full, 0);
        issue_qk(0);
        for (int j = 0; j < nt; ++j) {
            if (SBUF == 2 && j + 1 < nt) issue_qk(j + 1);
            const int st = j & 1, pb = j % PBUF;
            mbar_wait(&v_full[st], (j >> 1) & 1);
            mbar_wait(&p_full[pb], (j / PBUF) & 1);
            tc_fence_after();
            // (issuing QK(j+1) ahead of PV(j) here was measured: 91 -> 104 us at 4096 x 4126 keys -- PV(j) then completes later,
            // and with it the release of the V stage and of the P buffer the next softmax needs)
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk) {
                const uint64_t bd = umma_desc_sw128(sv + st * C::V_BYTES + (kk >> 2) * (DV * 128)) + 2 * (kk & 3);
                if constexpr (PTM) {       // 16 keys = 8 packed columns of the P buffer
                    umma_f16_ts_elect(tmem_base + C::TM_O, tmem_base + C::TM_P + pb * (BKV / 2) + kk * 8, bd, idesc_pv, (j | kk) != 0);
                } else {
                    const uint64_t ad = umma_desc_sw128(sp + pb * C::P_BYTES + (kk >> 2) * (128 * 128)) + 2 * (kk & 3);
                    umma_f16_elect(tmem_base + C::TM_O, ad, bd, idesc_pv, (j | kk) != 0);
                }
            }
            umma_commit_elect(bar_base + 8 * (7 + st));      // v_empty[st]
            umma_commit_elect(bar_base + 8 * (15 + pb));     // pv_done[pb]
            if (SBUF == 1 && j + 1 < nt) issue_qk(j + 1);
        }
    } else {
        // ---- softmax warps: one thread per query row.  Scores are consumed in 32-column chunks straight from TMEM.
        // Steady state is ONE pass: exponentiate against the running reference maximum m_used (exp2 domain) and only
        // when a row's tile maximum exceeds it by more than 2^8 (or on the first tile) take the slow path: fix the
        // reference, rescale O and l, and redo the tile.  P <= 256 fits fp16; l and O are fp32.
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const int wg = (warp - 2) >> 2;                       // 0 .. NWG-1
        constexpr int CW = BKV / NWG;                         // score columns per thread and tile
        const int cb = wg * CW;
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
        float* xch = reinterpret_cast<float*>(smem + C::OFF_XCH);   // [2][128]
        float m_used = -INFINITY, l = 0.f;
        const float sc = args.scale_log2;
        // O columns (16-wide chunks) this thread rescales / writes: split between the warpgroups
        auto owns = [&](int c) { return ((c >> 4) * NWG) / (DV / 16) == wg; };
        for (int j = 0; j < nt; ++j) {
            const int sb = j % SBUF, pb = j % PBUF;
            uint8_t* sP = smem + C::OFF_P + pb * C::P_BYTES;
            // columns of this tile visible to this row: the valid keys, under a causal mask only those up to the row's own index
            const int kvalid = CAUSAL ? min(args.nk - j * BKV, q0 + r + 1 - j * BKV) : args.nk - j * BKV;   // non-causal: >= 1
            const bool tail = CAUSAL || kvalid < BKV;
            const uint32_t ts = trow + C::TM_S + sb * BKV;
            mbar_wait(&s_full[sb], (j / SBUF) & 1);
            tc_fence_after();
            // (deferring this wait to the first P store of the tile was measured slower: 91 -> 103 us)
            if (j >= PBUF) {
                mbar_wait(&pv_done[pb], ((j / PBUF) - 1) & 1);      // P[pb] is free again (PBUF == 1: and O is up to date)
                tc_fence_after();
            }
            // exponentiate this thread's columns against reference `mref`, write P, return their sum; or (store = false)
            // only track the raw maximum
            auto pass = [&](auto tail_tag, auto store_tag, float mref, float& tmax) -> float {
                constexpr bool TAIL = decltype(tail_tag)::value;
                constexpr bool store = decltype(store_tag)::value;
                float rs = 0.f;
                const float nm = -mref;
                // (reading the tile in 16-column pieces with the next tcgen05.ld in flight behind the exponentials -- wait::ld tied to
                // the piece's registers -- was measured: 91.2 us either way at 4096 x 4126 keys; the 16 softmax warps per SM already
                // cover each other's TMEM latency)
#pragma unroll
                for (int c = 0; c < CW; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(ts + cb + c, v);
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        f[i] = __uint_as_float(v[i]);
                        if (TAIL && cb + c + i >= kvalid) f[i] = -INFINITY;
                        if (!store) tmax = fmaxf(tmax, f[i]);
                    }
                    [[maybe_unused]] uint32_t pk[16];      // PTM: the chunk's 32 probabilities as 16 packed half2
                    if (store) {
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            __half2 h[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int e0 = c8 * 8 + 2 * i;         // compile-time after unrolling
                                const float a0 = fmaf(f[e0], sc, nm), a1 = fmaf(f[e0 + 1], sc, nm);
                                constexpr int PD = POLY > 0 ? POLY : 1;
                                const float p0 = (POLY > 0 && e0 % PD == 0) ? ex2_poly(a0) : ex2_approx(a0);
                                const float p1 = (POLY > 0 && (e0 + 1) % PD == 0) ? ex2_poly(a1) : ex2_approx(a1);
                                rs += p0 + p1;
                                h[i] = __floats2half2_rn(p0, p1);
                            }
                            if constexpr (PTM) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) pk[c8 * 4 + i] = *reinterpret_cast<uint32_t*>(&h[i]);
                            } else {
                                const int g8 = ((cb + c) >> 3) + c8;       // 8-column group inside the tile
                                const int kc = g8 >> 3, u = g8 & 7;
                                *reinterpret_cast<uint4*>(sP + kc * (128 * 128) + r * 128 + ((u ^ (r & 7)) << 4)) =
                                    *reinterpret_cast<uint4*>(h);
                            }
                        }
                        if constexpr (PTM) tmem_st16(trow + C::TM_P + pb * (BKV / 2) + ((cb + c) >> 1), pk);
                    }
                }
                return rs;
            };
            auto exp_pass = [&](float mref) -> float {
                float unused = 0.f;
                return tail ? pass(BoolTag<true>{}, BoolTag<true>{}, mref, unused) : pass(BoolTag<false>{}, BoolTag<true>{}, mref, unused);
            };
            // optimistic pass against the current reference; every p >= 0, so a row sum below the fp16 range proves
            // that no single p overflowed -- otherwise (or on the first tile) fix the reference and redo the tile.
            // NWG == 2: the whole CTA votes, so both threads of a row always take the same path.
            float rs = 0.f;
            bool ok = false;
            if (j > 0) {
                rs = exp_pass(m_used);
                ok = rs <= 32768.0f;
            }
            const bool all_ok = NWG == 1 ? __all_sync(0xffffffffu, ok) : bar_red_and(1, 128 * NWG, ok);
            if (!all_ok) {
                float tmax = -INFINITY;
                if (tail) pass(BoolTag<true>{}, BoolTag<false>{}, 0.f, tmax);
                else pass(BoolTag<false>{}, BoolTag<false>{}, 0.f, tmax);
                if (NWG > 1) {                      // row maximum over the warpgroups' column ranges
                    xch[wg * 128 + r] = tmax;
                    named_bar_sync(2, 128 * NWG);
#pragma unroll
                    for (int w = 0; w < NWG; ++w) tmax = fmaxf(tmax, xch[w * 128 + r]);
                    named_bar_sync(2, 128 * NWG);   // xch is reused by the next exchange
                }
                const float mx = tmax * sc;
                const bool need = mx > m_used + 8.0f;
                const float m_new = need ? mx : m_used;
                if (j > 0 && __any_sync(0xffffffffu, need)) {
                    if (PBUF == 2) {      // O must hold every PV up to tile j-1 before it is rescaled
                        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
                        tc_fence_after();
                    }
                    const float f = need ? ex2_approx(m_used - m_new) : 1.0f;
                    l *= f;
#pragma unroll
                    for (int c = 0; c < DV; c += 16) {
                        if (!owns(c)) continue;
                        uint32_t v[16];
                        tmem_ld16(trow + C::TM_O + c, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
                        tmem_st16(trow + C::TM_O + c, v);
                    }
                    tmem_st_wait();
                }
                m_used = m_new;
                rs = exp_pass(m_used);
            }
            l += rs;
            if constexpr (PTM) tmem_st_wait();      // this thread's P columns are in tensor memory
            tc_fence_before();
            mbar_arrive(&s_empty[sb]);
            if constexpr (!PTM) fence_async_smem();
            mbar_arrive(&p_full[pb]);
        }
        mbar_wait(&pv_done[(nt - 1) % PBUF], ((nt - 1) / PBUF) & 1);
        tc_fence_after();
        if (NWG > 1) {                              // row sum over the warpgroups, fixed order
            xch[wg * 128 + r] = l;
            named_bar_sync(2, 128 * NWG);
            l = 0.f;
#pragma unroll
            for (int w = 0; w < NWG; ++w) l += xch[w * 128 + r];
        }
        const float inv = 1.0f / l;
        const bool valid = (q0 + r) < args.nq;
        __half* orow = args.out + ((size_t)b * args.nq + q0 + r) * args.ldo + head * args.dhead;
#pragma unroll
        for (int c = 0; c < DV; c += 16) {
            if (!owns(c)) continue;
            uint32_t v[16];
            tmem_ld16(trow + C::TM_O + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
                const int col = c + h8 * 8;
                if (valid && col < args.dhead) {
                    __half2 h[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        h[i] = __floats2half2_rn(__uint_as_float(v[h8 * 8 + 2 * i]) * inv,
                                                 __uint_as_float(v[h8 * 8 + 2 * i + 1]) * inv);
                    *reinterpret_cast<uint4*>(orow + col) = *reinterpret_cast<uint4*>(h);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TM_COLS>(tmem_base);
    }
}

template <int DPAD, int DV, int BKV, int SBUF, int PBUF, int MINB, int NWG = 1, int POLY = 0, bool PTM = false, bool CAUSAL = false>
static int attn_launch_variant(const AttnDeviceArgs& a, dim3 grid, cudaStream_t stream) {
    using C = AttnCfg<DPAD, DV, BKV, SBUF, PBUF, PTM>;
    static bool configured = false;
    if (!configured) {
        LTT_CUDA_OK(cudaFuncSetAttribute(attn_tc_kernel<DPAD, DV, BKV, SBUF, PBUF, MINB, NWG, POLY, PTM, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        configured = true;
        if (getenv("LTT_VERBOSE")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_tc_kernel<DPAD, DV, BKV, SBUF, PBUF, MINB, NWG, POLY, PTM, CAUSAL>, att_threads(NWG), C::TOTAL);
            fprintf(stderr, "[ltt] attn_tc_kernel<%d,%d,%d,%d,%d,%d>: %d B smem, %d CTA/SM\n", DPAD, DV, BKV, SBUF, PBUF, MINB, C::TOTAL, nb);
        }
    }
    LTT_CUDA_OK(launch_k(attn_tc_kernel<DPAD, DV, BKV, SBUF, PBUF, MINB, NWG, POLY, PTM, CAUSAL>, grid, dim3(att_threads(NWG)), C::TOTAL, stream, a));
    LTT_CUDA_OK(cudaGetLastError());
    return 0;
}

int attn_tc_launch(const AttnProblem& p, cudaStream_t stream) {
    AttnDeviceArgs a;
    memset(&a, 0, sizeof(a));
    const int rowlen = p.heads * p.dpad;
    int bkv, dv;
    static const int v40 = getenv("LTT_ATTN40") ? atoi(getenv("LTT_ATTN40")) : 0;      // d = 40 variant (experiments)
    if (p.dhead == 40 && p.dpad == 64) { bkv = (v40 == 1 || v40 == 2 || v40 == 8 || v40 == 9 || v40 == 10 || (v40 == 0 && p.nk <= 128)) ? 64 : 128; dv = 48; }
    else if (p.dhead == 80 && p.dpad == 128) { bkv = 128; dv = 80; }
    else if (p.dhead == 160 && p.dpad == 192) { bkv = 64; dv = 160; }
    else if (p.dhead == 64 && p.dpad == 64) { bkv = 128; dv = 64; }      // CLIP text tower (12 heads of 64)
    else if (p.dhead == 8 && p.dpad == 64) { bkv = 128; dv = 16; }     // tiny test / tiny-UNet heads
    else if (p.dhead == 16 && p.dpad == 64) { bkv = 128; dv = 16; }
    else {
        set_error("attention: unsupported head dim %d (dpad %d)", p.dhead, p.dpad);
        return -1;
    }
    if (p.causal && (p.nq != p.nk || p.dhead != 64)) {
        set_error("attention: the causal variant is built for 64-wide heads and nq == nk (d=%d, %d, %d)", p.dhead, p.nq, p.nk);
        return -1;
    }
    if (p.nk < 1 || p.nq < 1 || p.pitch_v % 8 != 0) {
        set_error("attention: bad sizes nq=%d nk=%d pitch_v=%d", p.nq, p.nk, p.pitch_v);
        return -1;
    }
    {
        uint64_t dims[3] = {(uint64_t)rowlen, (uint64_t)p.nq, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)rowlen, (uint64_t)rowlen * p.rows_q};
        uint32_t box[3] = {64, 128, 1};
        int rc = make_tmap_f16(&a.qmap, p.q, 3, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)rowlen, (uint64_t)p.nk, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)rowlen, (uint64_t)rowlen * p.rows_k};
        uint32_t box[3] = {64, (uint32_t)bkv, 1};
        int rc = make_tmap_f16(&a.kmap, p.k, 3, dims, str, box);
        if (rc) return rc;
    }
    {
        const int Cc = p.heads * p.dhead;
        uint64_t dims[3] = {(uint64_t)p.nk, (uint64_t)Cc, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.pitch_v, (uint64_t)p.pitch_v * Cc};
        uint32_t box[3] = {64, (uint32_t)dv, 1};
        int rc = make_tmap_f16(&a.vmap, p.vt, 3, dims, str, box);
        if (rc) return rc;
    }
    a.nq = p.nq; a.nk = p.nk; a.dhead = p.dhead; a.dpad = p.dpad;
    a.out = p.out; a.ldo = p.ldo;
    a.scale_log2 = p.scale * 1.4426950408889634f;
    dim3 grid((p.nq + 127) / 128, p.heads, p.B);
    static const int vwg = getenv("LTT_ATTN_WG") ? atoi(getenv("LTT_ATTN_WG")) : 2;     // softmax warpgroups for d = 80 / 160 (A/B)
    if (dv == 48) {
        if (v40 == 1) return attn_launch_variant<64, 48, 64, 1, 1, 3>(a, grid, stream);
        if (v40 == 2) return attn_launch_variant<64, 48, 64, 2, 2, 2>(a, grid, stream);
        if (v40 == 7) return attn_launch_variant<64, 48, 128, 1, 1, 2>(a, grid, stream);
        if (v40 == 8) return attn_launch_variant<64, 48, 64, 2, 2, 2, 2>(a, grid, stream);        // 64-key tiles, S and P double buffered, 2 warpgroups
        if (v40 == 9) return attn_launch_variant<64, 48, 64, 2, 2, 2, 2, 8>(a, grid, stream);     // ... + 1/8 software exp2
        if (v40 == 10) return attn_launch_variant<64, 48, 64, 2, 2, 2, 2, 8, true>(a, grid, stream);    // ... + P in tensor memory
        // default: two softmax warpgroups per CTA, two CTAs per SM (16 softmax warps per SM; measured alternatives at
        // 4096 x 4126 keys: 1 warpgroup 114 us, 2 warpgroups 103 us, 4 warpgroups 111 us, any 1-CTA/SM layout >= 129 us);
        // short key sets (the
        // 77-token text context) use 64-key tiles with S and P double buffered
        // P in tensor memory (PTM): 4096 x 4126 keys 97.3 -> 91.2 us, 77 keys 9.1 -> 8.1 us; LTT_ATTN_PTM=0: shared-memory P (A/B)
        static const int ptm = getenv("LTT_ATTN_PTM") ? atoi(getenv("LTT_ATTN_PTM")) : 1;
        if (bkv == 64) return ptm ? attn_launch_variant<64, 48, 64, 2, 2, 2, 2, 8, true>(a, grid, stream)
                                  : attn_launch_variant<64, 48, 64, 2, 2, 2, 2>(a, grid, stream);
        // software-exp2 share: every 8th exponential on the FMA / ALU pipes is the measured optimum (4096 x 4126 keys, B=2:
        // none 106.2 us, 1/16 105.5, 1/12 101.8, 1/8 97.6, 1/6 99.8, 1/4 101.5, 1/3 114.6, 1/2 122.6 -- beyond 1/8 the extra
        // nine instructions per element make the softmax warps issue bound); LTT_ATTN_POLY=0 switches it off (A/B)
        static const int poly = getenv("LTT_ATTN_POLY") ? atoi(getenv("LTT_ATTN_POLY")) : 8;
        if (ptm && poly == 8) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 8, true>(a, grid, stream);
        if (ptm) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 0, true>(a, grid, stream);
        if (poly == 2) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 2>(a, grid, stream);
        if (poly == 3) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 3>(a, grid, stream);
        if (poly == 4) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 4>(a, grid, stream);
        if (poly == 6) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 6>(a, grid, stream);
        if (poly == 8) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 8>(a, grid, stream);
        if (poly == 12) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 12>(a, grid, stream);
        if (poly == 16) return attn_launch_variant<64, 48, 128, 1, 1, 2, 2, 16>(a, grid, stream);
        return attn_launch_variant<64, 48, 128, 1, 1, 2, 2>(a, grid, stream);
    }
    // P in tensor memory for the wider heads too: d = 80 13.1 -> 12.3 us (1024 x 1054 keys), d = 160 7.3 -> 6.9 us
    static const int ptm_all = getenv("LTT_ATTN_PTM") ? atoi(getenv("LTT_ATTN_PTM")) : 1;
    // d = 64 (CLIP towers): the vision tower's 257-token sequences give 768+ CTAs of three KV tiles each -> two CTAs per SM
    // (one S and one P buffer: 256 TMEM columns) instead of one CTA with both double buffered; LTT_ATTN64=1: the latter (A/B)
    static const int v64 = getenv("LTT_ATTN64") ? atoi(getenv("LTT_ATTN64")) : 0;
    if (dv == 64 && p.causal) return attn_launch_variant<64, 64, 128, 2, 2, 1, 2, 0, true, true>(a, grid, stream);
    if (dv == 64) return v64 == 1 ? attn_launch_variant<64, 64, 128, 2, 2, 1, 2, 0, true>(a, grid, stream)
                                  : attn_launch_variant<64, 64, 128, 1, 1, 2, 2, 0, true>(a, grid, stream);
    if (dv == 80 && ptm_all) return attn_launch_variant<128, 80, 128, 2, 2, 1, 2, 0, true>(a, grid, stream);
    if (dv == 160 && ptm_all) return attn_launch_variant<192, 160, 64, 2, 2, 1, 2, 0, true>(a, grid, stream);
    if (dv == 80) return vwg == 2 ? attn_launch_variant<128, 80, 128, 2, 2, 1, 2>(a, grid, stream)
                                  : attn_launch_variant<128, 80, 128, 2, 2, 1>(a, grid, stream);
    if (dv == 160) return vwg == 2 ? attn_launch_variant<192, 160, 64, 2, 2, 1, 2>(a, grid, stream)
                                   : attn_launch_variant<192, 160, 64, 2, 2, 1>(a, grid, stream);
    return attn_launch_variant<64, 16, 128, 2, 2, 1>(a, grid, stream);
}

}  // namespace ltt
