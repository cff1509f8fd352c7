/* C-ABI of the B200-native layout-conditioned diffusion hot path (libltt_b200.so).
 *
 * The reference (LayoutLLM-T2I, /root/reference) is pure Python/PyTorch and has no FFI; the boundary this library
 * sits behind is the Python duck type described in SURVEY.md section 8(b).  Each entry point names the reference
 * interface it replaces.  Conventions: plain pointers and sizes only, every pointer is a DEVICE pointer unless it
 * says "host", the caller owns all I/O buffers, the library owns weights/workspaces, `stream` is a cudaStream_t
 * (pass torch.cuda.current_stream().cuda_stream), one handle per device, not thread safe, stream ordered.
 * Return value: 0 on success, negative on error (ltt_last_error() gives the message); nothing throws.
 *
 * fp16 = IEEE half.  "NHWC" = [B, H, W, C] with C contiguous; token tensors [B, N, C] are the same thing.
 */
#ifndef LTT_B200_H
#define LTT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ltt_model ltt_model;

enum { LTT_F16 = 0, LTT_F32 = 1 };
enum { LTT_ACT_NONE = 0, LTT_ACT_SILU = 1, LTT_ACT_GEGLU = 2 };

const char* ltt_last_error(void);
/* library / build identification: "ltt_b200 <version> sm_100a" */
const char* ltt_version(void);

/* ----------------------------------------------------------------------------------------------------------------
 * Model-level API  (replaces ldm.modules.diffusionmodules.openaimodel.UNetModel + ldm.models.diffusion.plms)
 * -------------------------------------------------------------------------------------------------------------- */

/* UNetModel.__init__ hyper-parameters (openaimodel.py:235-256; GLIGEN/configs/coco2014.yaml:8-30). */
typedef struct {
    int in_channels, out_channels, model_channels;
    int num_res_blocks;
    int n_levels;
    int channel_mult[8];
    int n_attn_res;
    int attention_resolutions[8];
    int num_heads;
    int context_dim;
    int grounding_in_dim, grounding_out_dim, fourier_freqs;
    int max_objs;      /* 30 (txt2img.py:173) */
} ltt_unet_config;

int ltt_create(const ltt_unet_config* cfg, int device, ltt_model** out);
void ltt_destroy(ltt_model* m);

/* model.load_state_dict (txt2img.py:106): one call per state_dict entry, reference key names and shapes
 * (SURVEY.md Appendix B); `data` is an fp32 DEVICE or HOST pointer (is_host != 0).  Unknown keys return -5. */
int ltt_load_param(ltt_model* m, const char* key, const float* data, const int64_t* shape, int ndim, int is_host);
/* Repack everything loaded so far into the device-native fp16 layouts; must precede the first forward.  Consumes
 * (frees) the fp32 staging copies of every matrix it repacked: after a weight change load the WHOLE state_dict again
 * (as model.load_state_dict does) before calling it again. */
int ltt_finalize(ltt_model* m);
/* UNetModel.restore_first_conv_from_SD (openaimodel.py:393-405): substitute input_blocks.0.0 (weight [Cm,4,3,3],
 * bias [Cm], fp32 device or host).  Permanent, like the reference, until ltt_clear_first_conv (a later
 * model.load_state_dict overwrites the swapped conv in the reference too). */
int ltt_set_first_conv(ltt_model* m, const float* weight, const float* bias, int is_host);
int ltt_clear_first_conv(ltt_model* m);

/* Per-image conditioning (everything UNetModel.forward derives from the non-x inputs, openaimodel.py:413-446):
 * context [B,77,ctx], relations [B,R,ctx], boxes [B,max_objs,4], masks [B,max_objs], pos_emb [B,max_objs,in_dim],
 * all fp32.  `n_grounded` leading batch elements use the given grounding, the rest use the null grounding input
 * (GroundingNetInput.get_null_input) -- that is how a [cond ; uncond] CFG batch is expressed.  Computes and caches
 * PositionNet tokens and all step-invariant K/V projections. */
int ltt_set_conditioning(ltt_model* m, const float* context, int ctx_len, const float* relations, int n_rel,
                         const float* boxes, const float* masks, const float* pos_emb, int B, int n_grounded,
                         int H, int W, void* stream);

/* UNetModel.forward (openaimodel.py:413-459) with the cached conditioning: x [B,4,H,W] fp32 NCHW, timesteps [B]
 * (fp32 values of the integer steps), gate scale as written by set_alpha_scale (txt2img.py:46-50);
 * eps_out [B,4,H,W] fp32 NCHW (values are fp16-rounded, as the autocast reference returns).  B, H, W are the shape
 * of x and must equal what ltt_set_conditioning cached (-1 otherwise: no out-of-bounds access on a stale cache). */
int ltt_unet_forward(ltt_model* m, const float* x, const float* timesteps, int B, int H, int W, float alpha_scale,
                     float* eps_out, void* stream);

/* PLMSSampler.plms_sampling (plms.py:64-163) for a batch whose conditioning was set with
 * ltt_set_conditioning(B = 2*Bimg, n_grounded = Bimg) when guidance != 1 (else B = Bimg):
 * x_inout [Bimg,4,H,W] fp32 start noise -> final latent.  host arrays of length S (index order of the sampler's
 * tables, i.e. ascending t): timesteps, alphas (ddim_alphas), alphas_prev, sqrt_one_minus_alphas; alpha_sched[S] =
 * per-iteration gate scale (alpha_generator, txt2img.py:59-93); when it hits 0 the SD first conv (if set via
 * sd_conv_weight/bias, may be NULL) is substituted as plms.py:86-87 does. */
int ltt_plms_sample(ltt_model* m, float* x_inout, int Bimg, int S, const int* timesteps_host,
                    const float* alphas_host, const float* alphas_prev_host, const float* sqrt_1m_alphas_host,
                    const float* alpha_sched_host, float guidance, const float* sd_conv_weight,
                    const float* sd_conv_bias, void* stream);

/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
int64_t ltt_launch_count(const ltt_model* m);

/* Per-kernel-class device timing for bench.py's roofline.  on = 1: eager launches, every launch of a class bracketed by
 * CUDA events on the launch stream.  on = 2: the brackets are external event-record nodes INSIDE the captured CUDA graphs
 * of the evaluation (the mode the timed path runs in); a bracket then holds the times of its graph's last replay and
 * ltt_profile_report weights it by the graph's replay count (calling ltt_profile_enable(m, 2) again restarts the
 * counts).  on = 0: off.  ltt_profile_report synchronises and returns, for class cls (0 = tcgen05 GEMM / implicit conv,
 * 1 = tcgen05 attention, 2 = GroupNorm, 3 = LayerNorm, 4 = whole UNet forward, 5 = relation fusion: box pooling, folded
 * relation attention, scatter + norm2), the summed event time [ms], the
 * algorithmic FLOPs (2*M*N*K; 4*nq*nk*d per head) and bytes, and the launch count since it was enabled. */
int ltt_profile_enable(ltt_model* m, int on);
int ltt_profile_report(ltt_model* m, int cls, double* ms, double* flops, double* bytes, int64_t* launches);

/* Debug taps (parity tests only): while a device buffer is registered, every ltt_unet_forward records named fp32
 * copies of intermediate activations ([rows, cols] = pixel/token rows x channels) into it.  Names are the reference's
 * module paths ("input_blocks.1.1", "input_blocks.1.1:attn1", ...).  buf == NULL switches recording off. */
int ltt_debug_set_taps(ltt_model* m, float* buf, int64_t capacity_elems);
int ltt_debug_tap_count(const ltt_model* m);
int ltt_debug_tap_info(const ltt_model* m, int idx, char* name, int name_cap, int64_t* offset, int64_t* rows,
                       int64_t* cols);

/* ----------------------------------------------------------------------------------------------------------------
 * VAE decode + image post-processing  (replaces ldm.models.autoencoder.AutoencoderKL.decode, autoencoder.py:40-44 ->
 * ldm.modules.diffusionmodules.model.Decoder.forward, model.py:538-568; and the callers' clamp / *255 / uint8 / HWC
 * conversion, txt2img.py:320-323, GLIGEN/interface.py:541-545)
 * -------------------------------------------------------------------------------------------------------------- */
typedef struct ltt_vae ltt_vae;

/* AutoencoderKL.__init__ (autoencoder.py:17-30) / Decoder.__init__ (model.py:462-536) hyper-parameters
 * (GLIGEN/configs/coco2014.yaml:33-52); attn_resolutions must be empty (only the mid attention block exists). */
typedef struct {
    int ch, out_ch;
    int n_levels;
    int ch_mult[8];
    int num_res_blocks;
    int z_channels, embed_dim;
    float scale_factor;
} ltt_vae_config;

int ltt_vae_create(const ltt_vae_config* cfg, int device, ltt_vae** out);
void ltt_vae_destroy(ltt_vae* v);
/* autoencoder.load_state_dict (txt2img.py:107): reference key names ("decoder.up.3.block.0.conv1.weight",
 * "post_quant_conv.bias", ...), fp32 device or host pointers; encoder.* / quant_conv.* keys are not needed. */
int ltt_vae_load_param(ltt_vae* v, const char* key, const float* data, const int64_t* shape, int ndim, int is_host);
int ltt_vae_finalize(ltt_vae* v);
/* decode(z): z [B, z_channels, h, w] fp32 NCHW latents (as PLMSSampler.sample returns them).  Outputs (either may be
 * NULL): img_f32 [B, out_ch, 8h, 8w] fp32 NCHW = the tensor AutoencoderKL.decode returns (fp16-rounded values, as under
 * autocast); img_u8 [B, 8h, 8w, out_ch] uint8 = uint8((clamp(img, -1, 1) * 0.5 + 0.5) * 255), the HWC image the callers
 * hand to PIL -- fused into the last convolution, ready for ONE pinned device-to-host copy.  h * w % 64 == 0. */
int ltt_vae_decode(ltt_vae* v, const float* z, int B, int h, int w, float* img_f32, uint8_t* img_u8, void* stream);
int64_t ltt_vae_launch_count(const ltt_vae* v);

/* ----------------------------------------------------------------------------------------------------------------
 * Conditioning prep: the CLIP text tower  (replaces the per-string calls of
 * ldm.modules.encoders.modules.FrozenCLIPEmbedder.forward / encode_one_token, modules.py:157-182 -- transformers
 * CLIPTextModel -- and of txt2img.get_clip_feature / extract_text_feat, txt2img.py:147-156,454-457 -- transformers
 * CLIPModel text tower -- by ONE batched pass over prompt, negative prompt, box phrases and relation phrases)
 * -------------------------------------------------------------------------------------------------------------- */
typedef struct ltt_clip ltt_clip;

/* transformers CLIPTextConfig fields (openai/clip-vit-large-patch14: 49408, 77, 768, 12, 12, 3072, 1e-5, quick_gelu,
 * projection 768, eos_token_id 2).  hidden / heads must be 64; hidden and ffn multiples of 64; act: 0 = quick_gelu. */
typedef struct {
    int vocab, max_pos, hidden, heads, layers, ffn;
    float eps;
    int act;
    int proj_dim;       /* 0: no text_projection (CLIPTextModel); > 0: CLIPModel.text_projection rows */
    int eos_token_id;   /* 2: pooled row = argmax(ids) (legacy config); otherwise first position holding this id */
} ltt_clip_config;

int ltt_clip_create(const ltt_clip_config* cfg, int device, ltt_clip** out);
void ltt_clip_destroy(ltt_clip* c);
/* state_dict keys of transformers CLIPTextModel ("text_model.embeddings.token_embedding.weight",
 * "text_model.encoder.layers.3.self_attn.q_proj.bias", ..., "text_model.final_layer_norm.weight") and, for proj_dim > 0,
 * CLIPModel's "text_projection.weight"; fp32 device or host pointers. */
int ltt_clip_load_param(ltt_clip* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_host);
/* packs the projection matrices to fp16 and releases their fp32 copies (a finalize after further ltt_clip_load_param calls
 * needs the whole state_dict loaded again, as load_state_dict does) */
int ltt_clip_finalize(ltt_clip* c);
/* ids [B, L] int32 on the device, L <= max_pos, rows laid out by CLIPTokenizer (<bos> ... <eos> padding); attention is
 * causal with NO padding mask (what FrozenCLIPEmbedder does; rows up to <eos> -- hence the pooled vector -- are the
 * same with one).  Outputs, fp32 device, each may be NULL: last_hidden [B, L, hidden] (= last_hidden_state, after
 * final_layer_norm), pooled [B, hidden] (= pooler_output), text_embeds [B, proj_dim] (= get_text_features). */
int ltt_clip_encode(ltt_clip* c, const int32_t* ids, int B, int L, float* last_hidden, float* pooled, float* text_embeds,
                    void* stream);
int64_t ltt_clip_launch_count(const ltt_clip* c);

/* ----------------------------------------------------------------------------------------------------------------
 * Reward path: CLIP vision tower + reward head  (replaces, in Reward.forward of
 * /root/reference/models/policy.py:106-123,139, the two eager `CLIPModel.get_image_features` calls -- transformers
 * CLIPVisionTransformer -- and the cosine / AestheticMLP (tools/aesthetic.py:9-31,52-57) arithmetic behind them; the text
 * features come from ltt_clip_encode's text_embeds)
 * -------------------------------------------------------------------------------------------------------------- */
typedef struct ltt_clip_vision ltt_clip_vision;

/* transformers CLIPVisionConfig fields (openai/clip-vit-large-patch14: 224, 14, 1024, 16, 24, 4096, 1e-5, quick_gelu,
 * projection 768).  hidden / heads must be 64; hidden and ffn multiples of 64; act: 0 = quick_gelu. */
typedef struct {
    int image_size, patch, hidden, heads, layers, ffn;
    float eps;
    int act;
    int proj_dim;       /* 0: no visual_projection (CLIPVisionModel); > 0: CLIPModel.visual_projection rows */
} ltt_clip_vision_config;

int ltt_clip_vision_create(const ltt_clip_vision_config* cfg, int device, ltt_clip_vision** out);
void ltt_clip_vision_destroy(ltt_clip_vision* c);
/* state_dict keys of transformers CLIPVisionModel / CLIPModel ("vision_model.embeddings.class_embedding",
 * "vision_model.embeddings.patch_embedding.weight", "vision_model.pre_layrnorm.weight" (sic), ...,
 * "vision_model.post_layernorm.bias", "visual_projection.weight"); fp32 device or host pointers. */
int ltt_clip_vision_load_param(ltt_clip_vision* c, const char* key, const float* data, const int64_t* shape, int ndim, int is_host);
int ltt_clip_vision_finalize(ltt_clip_vision* c);
/* pixel_values [B, 3, image_size, image_size] fp32 on the device, as CLIPProcessor returns them (resized, cropped,
 * normalised -- host-side preprocessing stays with the caller's processor).  Outputs, fp32 device, each may be NULL:
 * last_hidden [B, 1 + patches, hidden] (= last_hidden_state, before post_layernorm), pooled [B, hidden]
 * (= pooler_output), image_embeds [B, proj_dim] (= get_image_features). */
int ltt_clip_vision_encode(ltt_clip_vision* c, const float* pixel_values, int B, float* last_hidden, float* pooled,
                           float* image_embeds, void* stream);
/* The image preprocessing in front of the tower, on the device (replaces `self.processor(images=..., return_tensors="pt")`
 * of Reward.forward, models/policy.py:109-112 = transformers CLIPImageProcessor on PIL images: shortest edge -> image_size
 * with Pillow's BICUBIC resize, centre crop, * 1/255, (x - mean) / std): images [B, H, W, 3] uint8 on the device (the
 * layout ltt_vae_decode writes) -> pixel_values [B, 3, image_size, image_size] fp32.  Bit-exact against Pillow's
 * Resample.c (integer arithmetic, 8-bit rounding after each pass) and the float32 normalisation. */
int ltt_clip_vision_preprocess(ltt_clip_vision* c, const uint8_t* images, int B, int H, int W, const float* mean, const float* std_,
                               float* pixel_values, void* stream);
int64_t ltt_clip_vision_launch_count(const ltt_clip_vision* c);

/* Reward.forward models/policy.py:115-139 from the feature matrices: txt / pred / gt [B, D] fp32 (get_text_features of the
 * captions, get_image_features of the generated and the ground-truth images; D <= 1024);
 * clip = cos(txt, pred) + cos(gt, pred); aes = AestheticMLP(pred / |pred|) -- aes_w / aes_b: the five Linear layers
 * ("layers.0", ".2", ".4", ".6", ".7"), row-major [out, in] fp32 device pointers, aes_dims = {D, 1024, 128, 64, 16, 1};
 * reward[b] = clip + 0.1 aes + 10 miou[b] + 10 laysim[b]  (miou / laysim: the host-side layout terms, may be NULL = 0).
 * clip_reward / aes_reward [B] are optional outputs. */
int ltt_reward_head(const float* txt, const float* pred, const float* gt, int B, int D, const float* const* aes_w,
                    const float* const* aes_b, const int* aes_dims, const float* miou, const float* laysim, float* reward,
                    float* clip_reward, float* aes_reward, void* stream);

/* ----------------------------------------------------------------------------------------------------------------
 * Operator-level API (the same kernels, exposed one by one for parity tests and profiling)
 * -------------------------------------------------------------------------------------------------------------- */

/* nn.Linear / 1x1 conv:  out[M,N] = epilogue(a[M,K] . w[N,K]^T).  a, w fp16; K % 64 == 0, N % 8 == 0.
 * epilogue: y = acc + bias; act (SiLU, or GEGLU: w rows packed by ltt_op_pack_geglu, out has N/2 columns);
 * if has_gate y *= gate; if res: y += res.  (attention.py:38-65,108-112; openaimodel.py:172-194) */
int ltt_op_linear(const void* a, int M, int K, int lda, const void* w, int N, const float* bias, int act,
                  const void* res, int res_dtype, int ldr, float gate, int has_gate, void* out, int out_dtype, int ldo,
                  void* stream);
/* x = a . w1^T + b1 ; out2 = act(LayerNorm(x; gamma, beta, eps) . w2^T + b2)   (BasicTransformerBlock: `attn(norm(x))`,
 * `ff(norm(x))`, attention.py:394-402) with the LayerNorm FOLDED into the second GEMM: the first GEMM's epilogue leaves
 * per-row partial (sum, sum of squares), the second reads the raw fp16 rows x, uses w2 * gamma as weights and applies
 * rstd * (acc - mean * s_n) + c_n in its epilogue -- no normalised tensor is ever written.  a [M,K1] fp16, w1 [C,K1]
 * fp16, w2 [N,C] fp32 (packed inside), act2 0 or GEGLU; out1 [M,C] fp16, out2 [M, N or N/2] fp16. */
int ltt_op_linear_ln_linear(const void* a16, int M, int K1, const void* w1_16, const float* b1, int C, const float* gamma,
                            const float* beta, float eps, const float* w2_f32, const float* b2, int N, int act2, void* out1_16,
                            void* out2_16, void* stream);
/* [8C, K] fp32 GEGLU projection weight -> fp16 rows interleaved per 128-row tile (64 value rows, 64 gate rows) */
int ltt_op_pack_geglu(const float* w, int rows, int K, void* out_f16, void* stream);
/* OIHW fp32 3x3 conv weight [Cout, Cin, 3, 3] -> fp16 [Cout][tap][Cin] */
int ltt_op_pack_conv3x3(const float* w, int Cout, int Cin, void* out_f16, void* stream);
/* nn.Conv2d(C, N, 3, padding=1) on NHWC fp16 (openaimodel.py:155-185): out NHWC fp16 [B,H,W,N];
 * optional per-batch additive vector rowvec [B,N] fp16 (the ResBlock time-embedding add, :220-222) */
int ltt_op_conv3x3(const void* x, int B, int H, int W, int C, const void* w_packed, int N, const float* bias,
                   const void* rowvec, void* out, void* stream);
/* fused q/k/v projection writing attention-ready layouts: q,k [B, rows, heads*dpad] (head padded), vt [B, C, pitch] */
int ltt_op_qkv(const void* a, int B, int tokens, int C, const void* w_qkv, int heads, int dpad, void* q, int rows_q,
               void* k, int rows_k, void* vt, int pitch_v, void* stream);
/* softmax(q k^T scale) v  (attention.py:127-141,164-176), out [B, nq, heads*dhead] fp16 */
int ltt_op_attention(const void* q, int rows_q, const void* k, int rows_k, const void* vt, int pitch_v, int B,
                     int heads, int dhead, int dpad, int nq, int nk, float scale, void* out, int ldo, void* stream);
/* the same with a causal mask (key j visible to query i iff j <= i; n queries = n keys; dhead = dpad = 64): CLIP text tower self-attention
 * (transformers CLIPAttention with the causal mask of CLIPTextTransformer.forward) */
int ltt_op_attention_causal(const void* q, int rows_q, const void* k, int rows_k, const void* vt, int pitch_v, int B,
                            int heads, int dhead, int dpad, int n, float scale, void* out, int ldo, void* stream);
/* GroupNorm(32) [+ SiLU] on the channel concat of up to two NHWC fp16 tensors (util.py:211-229, attention.py:78) */
int ltt_op_groupnorm(const void* x0, int c0, const void* x1, int c1, int B, int HW, const float* gamma,
                     const float* beta, float eps, int silu, void* out, void* stream);
/* nn.LayerNorm(C) over rows (fp16 or fp32 in) -> fp16 and/or fp32 out */
int ltt_op_layernorm(const void* x, int x_dtype, int M, int C, const float* gamma, const float* beta, float eps,
                     void* out16, float* out32, void* stream);
/* RelationCrossAttention pooling / scatter pieces (attention.py:315-359) */
int ltt_op_rela_rects(const float* boxes, const float* masks, int B, int mo, int h, int w, int* rects, void* stream);
int ltt_op_rela_pool(const float* hid, const int* rects, int B, int mo, int h, int w, int C, void* feats16, void* stream);
int ltt_op_rela_scatter(const float* hid, const void* x16, const void* feats16, const int* rects, int nb_feats, int B,
                        int mo, int h, int w, int C, float* out, void* stream);
/* the same with the block's norm2 (attention.py:437, nn.LayerNorm eps 1e-5 on the fp32 stream) fused: ln16 = fp16 LN(out) */
int ltt_op_rela_scatter_ln(const float* hid, const void* x16, const void* feats16, const int* rects, int nb_feats, int B,
                           int mo, int h, int w, int C, float* out, const float* gamma, const float* beta, float eps,
                           void* ln16, void* stream);
/* Relation cross-attention with to_q / to_out folded into the step-invariant relation K / V (attention.py:315-359):
 * ltt_op_rela_fold: A[g,h*nrel+j,:] = scale * sum_{c in head h} Wq[c,:] k[g,j,c], Bm[g,h*nrel+j,:] = sum_{c in head h} Wo[:,c] v[g,j,c]
 *   (wq16 / wo16: [C, C] fp16 row-major nn.Linear weights, kv16: [G, nrel, 2C] fp16 = [k | v]; A16 / Bm16: [G, heads*nrel, C] fp16)
 * ltt_op_rela_attn_fused: per feature row f: f2 = f + gate * (softmax(LN1(f) . A^T) . Bm + bias), ln2 = LN2(f2)   (fp16 out) */
int ltt_op_rela_fold(const void* wq16, const void* wo16, const void* kv16, int G, int nrel, int heads, int d, float scale,
                     void* A16, void* Bm16, void* stream);
int ltt_op_rela_attn_fused(const void* feats16, int G, int rows_per_g, int C, int heads, int nrel, const void* A16,
                           const void* Bm16, const float* bias, float gate, const float* g1, const float* b1, const float* g2,
                           const float* b2, float eps, void* feats2_16, void* ln2_16, void* stream);
/* 30 x 10 relation cross-attention core (warp shuffles) */
int ltt_op_small_attention(const void* q, const void* k, const void* v, int B, int nq, int nk, int heads, int d,
                           float scale, void* out, void* stream);
/* PositionNet input rows [rows, in_dim + 8*nfreq] fp16 (text_grounding_net.py:26-41, util.py:12-26) */
int ltt_op_posnet_input(const float* boxes, const float* masks, const float* emb, const float* null_txt,
                        const float* null_pos, int rows, int in_dim, int nfreq, void* out16, void* stream);
/* timestep_embedding (util.py:161-181) -> fp16 [B, dim] */
int ltt_op_timestep_embedding(const float* t, int B, int dim, void* out16, void* stream);
/* CFG combine + PLMS update (plms.py:110-163); see small_ops.cu plms_update_kernel for `mode` */
int ltt_op_plms_update(const float* eps_c, const float* eps_u, float guidance, int use_cfg, int mode, const float* x,
                       float* e_t_out, const float* e_first, const float* old1, const float* old2, const float* old3,
                       float a_t, float a_prev, float sqrt_1m_at, float* x_out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTT_B200_H */
