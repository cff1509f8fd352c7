// Host launchers of the CUDA-core kernels in small_ops.cu (see that file for semantics and reference citations).
#pragma once
#include <algorithm>

#include "ltt_kernels.h"

namespace ltt {

int gn_stats_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                    double* stats, cudaStream_t st);
int gn_apply_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                    const double* stats, const float* gamma, const float* beta, float eps, int silu, __half* out,
                    cudaStream_t st);
int groupnorm_fused_launch(const __half* x0, int c0, int ld0, const __half* x1, int c1, int ld1, int B, int HW, int groups,
                           const float* gamma, const float* beta, float eps, int silu, __half* out, cudaStream_t st);
int layernorm_launch(const void* x, int x_dtype, int M, int C, const float* gamma, const float* beta, float eps,
                     __half* out16, float* out32, cudaStream_t st, float2* stats_out = nullptr);
int conv_in_launch(const float* x, const float* w, const float* bias, int B, int Cin, int H, int W, int Cout,
                   __half* out, cudaStream_t st);
int conv_out_launch(const __half* x, const __half* w, const float* bias, int B, int H, int W, int C, int Cout,
                    float* out, cudaStream_t st);
int upsample2x_launch(const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st);
int im2col_s2_launch(const __half* in, __half* out, int B, int H, int W, int C, cudaStream_t st);
int timestep_embed_launch(const float* t, int B, int dim, __half* out, cudaStream_t st);
int posnet_input_launch(const float* boxes, const float* masks, const float* emb, const float* null_txt,
                        const float* null_pos, int rows, int in_dim, int nfreq, __half* out, cudaStream_t st);
int rela_rects_launch(const float* boxes, const float* masks, int B, int mo, int h, int w, int* rects, cudaStream_t st);
// row statistics for kernels that apply a LayerNorm on the fly: slots == -1: p[row * ld] = (mean, rstd); slots > 0: partial
// (sum, sum of squares) over K columns in p[row * ld + 0 .. slots) (left by a GEMM epilogue, see gemm_tc.cu LayerNorm fold)
struct RowStatSrc {
    const float2* p = nullptr;
    int ld = 1, slots = -1, K = 0;
    float eps = 1e-5f;
};
int rela_pool_launch(const float* hid, const __half* x16, const RowStatSrc& stats, const float* gamma, const float* beta,
                     const int* rects, int B, int mo, int h, int w, int C, __half* feats, cudaStream_t st);
int rela_scatter_launch(const float* hid, const RowStatSrc& stats, const float* gamma3, const float* beta3, const __half* x,
                        const __half* feats, const int* rects, int nb_feats, int B, int mo, int h, int w, int C, float* out,
                        const float* gamma, const float* beta, float eps, __half* ln16, cudaStream_t st);
int rela_fold_launch(const __half* wq, const __half* wo, const __half* kv, int G, int nrel, int heads, int d, float scale,
                     __half* A, __half* Bm, cudaStream_t st);
int rela_attn_fused_launch(const __half* feats, int G, int rows_per_g, int C, int heads, int nrel, const __half* A,
                           const __half* Bm, const float* bias, float gate, const float* g1, const float* b1,
                           const float* g2, const float* b2, float eps, float* scratch, int* tickets, __half* feats2,
                           __half* ln2out, cudaStream_t st);
int small_attn_launch(const __half* q, int ldq, const __half* k, const __half* v, int ldkv, int B, int nq, int nk,
                      int heads, int d, float scale, __half* out, cudaStream_t st);
int cast_f32_f16_launch(const float* in, __half* out, size_t n, cudaStream_t st);
int cast_f16_f32_launch(const __half* in, float* out, size_t n, cudaStream_t st);
int copy2d_launch(const __half* src, size_t sb, int sld, __half* dst, size_t db, int dld, int B, int rows, int cols,
                  cudaStream_t st);
int plms_update_launch(const float* eps_c, const float* eps_u, float guidance, int use_cfg, int mode, const float* x,
                       float* e_t_out, const float* e_first, const float* old1, const float* old2, const float* old3,
                       float a_t, float a_prev, float sqrt_1m_at, float* x_out, size_t n, cudaStream_t st);

int plms_prep_launch(const float* x, float* x_in, size_t n, int copies, float* t_in, int B, float tval, const __half* ev_src,
                     __half* ev_dst, int ev_len, cudaStream_t st);
int fill_tvals_launch(const float* host_vals, int n, float* out, cudaStream_t st);

int pack_conv_launch(const float* w, int O, int Cin, int taps, int cstart, int Cs, __half* dst, int Kdst, int koff,
                     cudaStream_t st);
int pack_rows_launch(const float* w, int rows, int K, __half* dst, int row_off, int geglu, cudaStream_t st,
                     const float* colscale = nullptr);
int ln_fold_vectors_launch(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K, float* s_out,
                           float* c_out, cudaStream_t st);

}  // namespace ltt
