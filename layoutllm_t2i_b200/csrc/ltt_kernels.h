// Internal launcher interface of the sm_100a kernels (host side).  Everything here works on raw device pointers;
// activations are NHWC fp16 ("pixel/token rows x channels"), accumulators fp32.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ltt {

// ---------------------------------------------------------------------------------------------------------------
// error plumbing: kernels never throw across the C-ABI; launchers return 0 / negative and record a message.
void set_error(const char* fmt, ...);
const char* last_error();
#define LTT_CUDA_OK(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ltt::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -2;                                                                            \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Kernel launch with programmatic dependent launch enabled (see ltt_ptx.cuh pdl_wait); LTT_NO_PDL=1 turns it off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA descriptor helpers (cuTensorMapEncodeTiled resolved at run time through cudaGetDriverEntryPoint).
// fp16 tensor, up to 4 dims (dim0 innermost, contiguous), strides in ELEMENTS for dims 1..rank-1, 128-B swizzle.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                  const uint32_t* box);

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core GEMM / implicit-GEMM convolution:   D[m, n] = sum_k A[m, k] * Wp[n, k]
//   rows m = pixels (b, y, x) of an NHWC activation (or plain rows when H == 1),
//   A is gathered from up to 3 sources; a source with taps == 9 contributes the 3x3 neighbourhood (zero padded,
//   stride 1), one with taps == 1 its own pixel.  Wp is [N, Ktot] fp16, K ordered source-major, tap, channel.
enum : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GEGLU = 2, ACT_QUICKGELU = 3 };   // QUICKGELU: x * sigmoid(1.702 x) (CLIP text MLP)
enum : int { OUT_ROWMAJOR = 0, OUT_QKV = 1 };
enum : int { DT_F16 = 0, DT_F32 = 1 };

struct GemmSrc {
    const __half* ptr;   // NHWC base (channel 0 of the source)
    int channels;        // channels taken from this source (multiple of 64)
    int ld;              // channel pitch of the tensor in elements (>= channels, multiple of 8)
    int taps;            // 1 or 9
};

struct GemmEpilogue {
    const float* bias = nullptr;      // [N] (for GEGLU: packed order, see repack)
    const __half* rowvec = nullptr;   // [B, ld_rowvec] added per batch element (time embedding), fp16
    int ld_rowvec = 0;
    int act = ACT_NONE;
    float gate = 1.0f;                // out = res + gate * y   (gate applied only when has_gate)
    int has_gate = 0;
    const void* res = nullptr;        // residual [M, ldr]
    int res_dtype = DT_F16;
    int ldr = 0;
    void* out = nullptr;              // [M, ldo]
    int out_dtype = DT_F16;
    int ldo = 0;
    int out_mode = OUT_ROWMAJOR;
    // OUT_QKV: columns [0,C) -> q, [C,2C) -> k (head-padded rows), [2C,3C) -> v transposed
    __half* q = nullptr;  // [B, rows_q, heads*dpad]
    __half* k = nullptr;  // [B, rows_k, heads*dpad]
    __half* vt = nullptr; // [B, C, pitch_v]
    int C = 0, dhead = 0, dpad = 0, rows_q = 0, rows_k = 0, pitch_v = 0, tokens = 0;  // tokens = rows per batch elem
    int qkv_base = 0;                 // first C-wide column block is q (0), k (1) or v (2)
    // ---- LayerNorm folded into the CONSUMER GEMM (gemm_tc.cu, "LayerNorm fold").
    // Producer side: every epilogue thread leaves the partial (sum, sum of squares) of the fp16 values it stored for its
    // row in slot (n_tile * splits + k_rank) * EPI_WGS + warpgroup of stats_out[row * stats_ld + slot]; the launcher
    // reports the number of slots used.  out16: additional fp16 copy of an fp32 output (the LayerNorm input of the
    // fp32 residual stream reaches the consumer GEMM as fp16 rows), same pitch as `out`.
    float2* stats_out = nullptr;
    int stats_ld = 0;
    __half* out16 = nullptr;
    // Consumer side: A holds the RAW rows x, the packed weights are W' = W * gamma (column scale) and
    //   y[m, n] = rstd_m * (acc[m, n] - mean_m * ln_s[n]) + bias[n],   ln_s[n] = sum_k W'[n, k],  bias[n] = b[n] + sum_k beta[k] W[n, k]
    // with (mean, rstd) of row m over its ln_K columns from ln_slots partial sums (ln_slots > 0) or stored directly as
    // (mean, rstd) (ln_slots == -1).
    const float2* ln_stats = nullptr;
    int ln_slots = 0, ln_ld = 0, ln_K = 0;
    const float* ln_s = nullptr;
    float ln_eps = 1e-5f;
};

struct GemmProblem {
    int B, H, W;          // row geometry: M = B*H*W  (plain matrices: B = 1, H = 1, W = M)
    int N;                // output columns = rows of Wp (GEGLU: N is the packed 8C width; out has N/2 columns)
    int nsrc;
    GemmSrc src[3];
    const __half* w;      // packed weights [N, Ktot]
    int Ktot;             // = sum taps*channels
    int w_static;         // 1: `w` is never written by a kernel still in flight on the stream (model weights)
    GemmEpilogue epi;
    // host only: tiling forced by the autotuner (model.cu); force_bn == 0 -> the launcher's cycle model picks.
    // An illegal forced tiling returns GEMM_ILLEGAL_TILING without launching.
    int force_bn = 0, force_splits = 1, force_pair = 0;
};
constexpr int GEMM_ILLEGAL_TILING = -9;

// (split-K partial sums meet in distributed shared memory of a thread-block cluster: no global workspace)
// stats_slots (optional): receives the number of statistics slots per row the launch writes when epi.stats_out is set.
// chosen (optional): receives {tile width BN, split-K cluster size, CTA-pair flag} of the tiling that was launched.
int gemm_tc_launch(const GemmProblem& p, int num_sms, cudaStream_t stream, int* stats_slots = nullptr, int* chosen = nullptr);
constexpr int GEMM_STATS_LD = 128;      // slots per row every stats buffer provides (launches are tiled to fit)

// ---------------------------------------------------------------------------------------------------------------
// Fused attention core:  O[b, i, h*d:(h+1)*d] = softmax_j(q_i . k_j * scale) v_j   (flash-style, S/O in TMEM)
struct AttnProblem {
    int B, heads, dhead, dpad;
    int nq, nk;               // valid query / key rows per batch element
    const __half* q;          // [B, rows_q, heads*dpad]
    int rows_q;
    const __half* k;          // [B, rows_k, heads*dpad]
    int rows_k;
    const __half* vt;         // [B, heads*dhead, pitch_v]  (keys contiguous)
    int pitch_v;
    __half* out;              // [B, nq, ldo]
    int ldo;
    float scale;
    int causal = 0;           // 1: key j is visible to query i only when j <= i (CLIP text tower); needs nq == nk
};
int attn_tc_launch(const AttnProblem& p, cudaStream_t stream);

}  // namespace ltt
