/* clover_b200 -- C ABI of the B200-native (sm_100a) Clover hot-path kernels.
 *
 * This is the drop-in boundary of the project: the reference (LeeYN-43/Clover, an mmaction2 fork)
 * is pure Python/PyTorch and has NO FFI of its own, so each entry point below
 * names the reference Python code whose arithmetic it replaces (file:line under the reference
 * tree).  The Python host side (clover_b200/ops.py) binds these with ctypes; INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + explicit sizes/pitches (in ELEMENTS) + a cudaStream_t passed as void*;
 *   - every function is stream-ordered, re-entrant, allocates nothing persistent, never
 *     synchronises the device and frees nothing: the caller owns all memory;
 *   - return 0 on success; otherwise non-zero and clv_last_error() (thread-local) describes it;
 *   - "bf16" buffers hold __nv_bfloat16, "f32" buffers float; *_is_bf16 flags select per operand;
 *   - no CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef CLOVER_B200_H_
#define CLOVER_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* clv_last_error(void);
int clv_version(void);
/* Number of kernel launches issued by this library in the calling process (gpu_launches claim). */
long long clv_launch_count(void);
/* Experiment knobs for tools / tests ("gemm_bn256_min_units", "w7_pipe", "w7_bwd2", "w7_dbias_acc", "w7_dbias_acc_min_mb", "w7_fwd2", "gemm_tma_store");
 * value -1 restores the built-in default.  The library reads no environment variables.  Returns non-zero for unknown names. */
int clv_set_tunable(const char* name, long long value);

/* Window geometry of one Swin stage call: real extent (B,D,H,W), clamped window and shift as
 * returned by get_window_size (swin_transformer_3d.py:302-315).  Padding to window multiples
 * (:452-456) is derived inside.  Encodes fused roll(-shift)+window_partition (:460,:271-283)
 * and its inverse window_reverse+roll(+shift) (:471-474) in closed form. */
typedef struct {
  int B, D, H, W;
  int wd, wh, ww;
  int sd, sh, sw;
} clv_window_geom_t;

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA).  D[M,N] = epilogue(sum_k A[m,k] B[n,k]).
 *   A: K-major  -> row-major [M, K] with pitch lda;  MN-major -> stored [K, M] with pitch lda.
 *   B: K-major  -> row-major [N, K] (an nn.Linear weight);  MN-major -> stored [K, N].
 * Replaces every nn.Linear forward/backward on the path: qkv/proj (swin_transformer_3d.py:376,398),
 * Mlp fc1/fc2 (:263-266), PatchMerging.reduction (:542), PatchEmbed3D conv-as-GEMM (:681), HF BERT
 * dense layers (call sites bert_from_hugface.py:30, cross_transformer.py:67,110), head projections
 * (heads/ssl_head.py:50-69,181-185,257-260; mlm_itm_head.py:38-41; qa_head.py:59-65).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* bias;          /* [N] or NULL */
  const void* residual;       /* [M(out rows), N] or NULL, added last */
  int residual_is_bf16;
  long long ld_residual;
  void* out;                  /* required */
  int out_is_bf16;
  long long ld_out;
  void* out_pre;              /* bf16 pre-activation copy (act == 1) or NULL */
  long long ld_pre;
  const void* gelu_pre;       /* bf16 [M,N]: multiply result by GELU'(gelu_pre) (fc2 dgrad) or NULL */
  long long ld_gelu_pre;
  int act;                    /* 0 none, 1 exact erf GELU */
  int scale_cols;             /* columns [0,scale_cols) are multiplied by `scale` after the bias (q * hd^-0.5, :379) */
  float scale;
  const clv_window_geom_t* window;  /* non-NULL: GEMM row (window order) -> spatial output/residual row (:471-479) */
  int k_splits;               /* >1: split K across CTAs, fp32 atomic accumulation (weight gradients) */
  int accumulate;             /* 1: add into `out` (fp32) instead of overwriting */
  const float* row_scale;     /* [ceil(M / row_scale_rows)] or NULL: (acc + bias) of GEMM row m is multiplied by
                                 row_scale[m / row_scale_rows] before the residual add -- the per-sample keep/(1-p)
                                 factor of timm DropPath (x + drop_path(branch), swin_transformer_3d.py:499,503) */
  long long row_scale_rows;
  float* rowsum;              /* [M] fp32 or NULL: rowsum[m] += sum_k A[m, k] (ACCUMULATED with atomics; caller zero-fills).  With
                                 A = dY^T (a weight gradient dW = dY^T X) this is the nn.Linear bias gradient, computed by one extra
                                 tensor-core instruction per K step instead of a separate column-sum pass over dY */
} clv_gemm_epilogue_t;

int clv_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                  int M, int N, int K, const clv_gemm_epilogue_t* epilogue, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm family.  Replaces nn.LayerNorm at swin_transformer_3d.py:450,483,541,685,238 (eps 1e-5),
 * cross_transformer.py:98 (eps 1e-5), HF BERT LayerNorms (eps 1e-12), head LNs, together with the
 * layout copies around them (pad/roll/window_partition :456-466; PatchMerging slices+cat :535-539;
 * mask-token blend :222-230; fusion positional adds + concat cross_transformer.py:84-108).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x; int x_is_bf16; long long ld_x;
  const float* gamma; const float* beta; float eps;
  float* mean; float* rstd;          /* [rows] fp32; written by fwd, read by bwd (may be NULL in fwd) */
  long long rows; int C;             /* OUTPUT rows, normalised width */
  const clv_window_geom_t* window;   /* gather: output row r (window order) <- spatial row; padding -> 0 */
  int merge_B, merge_D, merge_H, merge_W, merge_C;   /* merge_C > 0: 2x2 patch-merging gather, C == 4*merge_C */
  const float* add0;                 /* [C] added before normalising */
  const float* add1; int div1, mod1; /* [mod1, C] row (r / div1) % mod1 */
  const float* add2; int div2, mod2;
  long long group_rows, group_stride, row_offset;    /* group_rows > 0: out row = (r/group_rows)*group_stride + r%group_rows + row_offset */
  const long long* blend_mask;       /* (B, mh, mw) int64 0/1: y = y*(1-w) + token*w after the LN */
  const float* blend_token;          /* [C] */
  int blend_D, blend_H, blend_W, blend_mh, blend_mw;
  const long long* row_index;        /* plain mode: source row = row_index[r] (HF BertEmbeddings word lookup) */
} clv_ln_desc_t;

int clv_layernorm_fwd(const clv_ln_desc_t* desc, void* y, int y_is_bf16, long long ld_y, void* stream);

typedef struct {
  const void* dy; int dy_is_bf16; long long ld_dy;
  float* dx; long long ld_dx;                  /* fp32 at the source rows (NULL: parameter grads only) */
  const float* dres; long long ld_dres;        /* optional residual-stream gradient added into dx */
  void* dx_copy; int dx_copy_is_bf16; long long ld_copy;   /* optional copy of the final dx */
  const clv_window_geom_t* copy_window;        /* copy written in window order of this geometry */
  float* dgamma; float* dbeta;                 /* [C] fp32, ACCUMULATED (caller zero-fills) */
  float* dtoken;                               /* [C] d(mask_token), blend only */
  int dx_dense;                                /* write dx at row r, not at the (gathered) source row */
  /* merge gather at the Swin widths (merge_C = 128 / 256 / 512, fp32 x, bf16 dy) only: the bf16 dx_copy and dxsum carry the
   * per-sample DropPath factor copy_scale[source row / copy_scale_rows] of the block that produced x (like clv_lnr_bwd_t);
   * dxsum [merge_C] fp32 ACCUMULATES the column sums of that scaled dx = the bias gradient of the block's fc2 */
  float* dxsum; const float* copy_scale; long long copy_scale_rows;
} clv_ln_bwd_t;

int clv_layernorm_bwd(const clv_ln_desc_t* desc, const clv_ln_bwd_t* bwd, void* stream);

/* Row-mapped LayerNorm on dense contiguous rows [rows, C] (pitch == C): the lean hot path of
 * SwinTransformerBlock3D's norm1 / norm2 (swin_transformer_3d.py:450,483) and of every plain nn.LayerNorm.
 * Statistics are indexed by the SOURCE row s.  row_map (int32 [map_period], one clip's permutation) sends
 * source row s to m(s) = (s / map_period) * map_period + row_map[s % map_period]: the fused
 * roll(-shift) + window_partition of :456-466 (its inverse is window_reverse + roll back, :471-479), built on
 * the host from the closed forms.  Unpadded geometries only; clv_lnr_supported(C) tells whether the width
 * has a specialisation (otherwise use clv_layernorm_*). */
typedef struct {
  const void* x; int x_is_bf16;
  const float* gamma; const float* beta; float eps;
  float* mean; float* rstd;          /* [rows] fp32 by source row; written by fwd (may be NULL), read by bwd */
  long long rows; int C;
  const int* row_map; int map_period;
  /* optional SimMIM mask-token blend after the patch-embed LayerNorm (swin_transformer_3d.py:222-230):
   * y[s] = LN(x[s]) * (1 - w[s]) + token * w[s]; backward: d(LN out) = dy * (1 - w), dtoken += sum_s dy[s] * w[s] */
  const float* row_blend;            /* [rows] fp32 weights w (0 / 1 from v_token_mask), or NULL */
  const float* blend_token;          /* [C] fp32 */
} clv_lnr_desc_t;

typedef struct {
  const void* dy; int dy_is_bf16;
  int dy_mapped;                     /* dy of source row s lives at row m(s) (norm1: gradient in window order) */
  const float* dres;                 /* fp32 [rows, C] residual-stream gradient added to dx, or NULL (may alias dx) */
  float* dx;                         /* fp32 [rows, C] or NULL */
  void* dx_bf16; int dx_bf16_mapped; /* optional bf16 copy of dx, at row m(s) when mapped (norm2: proj operand) */
  float* dgamma; float* dbeta;       /* [C] fp32, ACCUMULATED (caller zero-fills); both or neither */
  float* dxsum;                      /* [C] fp32, ACCUMULATED column sums of dx, or NULL: the bias gradient of the nn.Linear
                                        whose output gradient this dx is (fc2 / proj of the Swin block); fp32 x + bf16 dy only */
  const float* copy_scale;           /* NULL, or per-sample DropPath factors: the bf16 copy and dxsum (NOT dx) are multiplied by
                                        copy_scale[s / copy_scale_rows] -- they are the gradient of a residual branch that the
                                        forward pass scaled by the same factor (swin_transformer_3d.py:499,503) */
  long long copy_scale_rows;
  float* dtoken;                     /* [C] fp32, ACCUMULATED gradient of blend_token (row_blend set; excludes dxsum) */
} clv_lnr_bwd_t;

int clv_lnr_supported(int C);
int clv_lnr_fwd(const clv_lnr_desc_t* desc, void* y, int y_is_bf16, int y_mapped, void* stream);
int clv_lnr_bwd(const clv_lnr_desc_t* desc, const clv_lnr_bwd_t* bwd, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention core on packed qkv rows. qkv: bf16 [batch*seq, 3*heads*hd] laid out [3][heads][hd]
 * per row (the reshape of swin_transformer_3d.py:376 and of a fused BERT q|k|v dense), q already
 * scaled.  out: bf16 [batch*seq, heads*hd]; lse: fp32 [batch, heads, seq] (saved for backward).
 *   bias_table (+ rel_code)  : relative-position bias table[(code_i - code_j + code_off), head]
 *                              (:341-359, :382-385), table fp32 [table_len, heads]
 *   region (+ nwin)          : shift mask, 0 if region[win][i] == region[win][j] else -100
 *                              (:388-390, compute_mask :548-562); win = batch % nwin
 *   key_mask                 : fp32 additive per key [batch, seq] ((1-m)*-10000, HF BERT 4.6.1)
 * Replaces WindowAttention3D.forward :379-397 and HF BertSelfAttention's softmax(QK^T/8+M)V.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int batch, seq, heads, head_dim;       /* head_dim 32 or 64; seq <= 512 */
  const float* bias_table; int table_len;
  const int* rel_code; int code_off;     /* [seq] */
  const int* region; int nwin;           /* [nwin, seq] or NULL */
  const float* key_mask;                 /* [batch, seq] or NULL */
  /* attention-probability dropout (HF BertSelfAttention.dropout, attention_probs_dropout_prob): P is multiplied by
   * keep/(1-p) after the softmax, keep = clv_keep_mask stream element drop_offset + ((b*heads+h)*seq+i)*seq+j.
   * drop_p == 0 disables it.  Only clv_attention_fwd / clv_attention_bwd (the BERT kernels) support it. */
  float drop_p;
  unsigned long long drop_seed, drop_offset;
} clv_attn_desc_t;

int clv_attention_fwd(const clv_attn_desc_t* desc, const void* qkv, void* out, float* lse, void* stream);
/* Same contract on tcgen05 / TMEM / TMA: head_dim 32, 33 <= seq <= 416, no key_mask (the Video Swin windows). */
int clv_attention_fwd_tc(const clv_attn_desc_t* desc, const void* qkv, void* out, float* lse, void* stream);
/* dqkv: bf16 same layout as qkv (dq already multiplied by q_scale so it is d/d(unscaled q));
 * dbias_table: fp32 [table_len, heads] ACCUMULATED (may be NULL). */
int clv_attention_bwd(const clv_attn_desc_t* desc, const void* qkv, const void* out, const void* dout,
                      const float* lse, void* dqkv, float q_scale, float* dbias_table,
                      float* dsum_workspace /* fp32 [batch*heads*seq] */, void* stream);
/* The BERT / fusion attention (head_dim 64, optional key_mask, optional attention-probability dropout; no bias table / region)
 * on tcgen05 / TMEM / TMA: HF BertSelfAttention (transformers 4.6.1) at the reference's call sites bert_from_hugface.py:30 and
 * cross_transformer.py:109-110.  Same contract as clv_attention_fwd / clv_attention_bwd for 1 <= seq <= 448
 * (clv_attention_tc64_supported); the backward recomputes the probabilities in a dK|dV and a dQ kernel. */
int clv_attention_tc64_supported(int seq);
int clv_attention_fwd_tc64(const clv_attn_desc_t* desc, const void* qkv, void* out, float* lse, void* stream);
int clv_attention_bwd_tc64(const clv_attn_desc_t* desc, const void* qkv, const void* out, const void* dout, const float* lse,
                           void* dqkv, float q_scale, float* dsum_workspace /* fp32 [batch*heads*seq] */, void* stream);
/* probs_mean[b, i, j] = mean over heads of softmax_j(q_i . k_j + bias/mask terms): the `attentions[-1].mean(dim=1)`
 * that CloverFinetune.forward_test returns for video QA (multimodal_transformer_finetune.py:192; HF BertEncoder
 * output_attentions).  Evaluation only (no dropout).  out: fp32 [batch, seq, seq]. */
int clv_attention_probs_mean(const clv_attn_desc_t* desc, const void* qkv, float* probs_mean, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound helpers.
 * ------------------------------------------------------------------------------------------- */
/* dst[i] = scale * src[i] with dtype conversion (fp32 master weights -> bf16 operand caches, the
 * model.half() of core/hooks/fp16_utils.py:215-239; gradient-stream casts).  n % 4 == 0. */
int clv_cast(const void* src, int src_is_bf16, void* dst, int dst_is_bf16, long long n, float scale, void* stream);

/* Exact erf GELU (nn.GELU() in heads/ssl_head.py:53,67,183,259 and qa_head.py:14,63).  dy == NULL: y = GELU(x);
 * otherwise y = dy * GELU'(x).  n % 4 == 0. */
int clv_gelu(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, void* y, int y_is_bf16, long long n, void* stream);

/* nn.Tanh of ITMHead.itm_projector (heads/mlm_itm_head.py:67-70).  dy == NULL: y = tanh(x); otherwise x holds the saved
 * tanh OUTPUT t and y = dy * (1 - t^2).  n % 4 == 0. */
int clv_tanh(const void* x, int x_is_bf16, const void* dy, int dy_is_bf16, void* y, int y_is_bf16, long long n, void* stream);

/* PatchEmbed3D's Conv3d(kernel == stride) as a patch matrix (swin_transformer_3d.py:665,671-681):
 * x fp32 (B,Cin,F,H,W) -> bf16 [B*D*Hp*Wp, Cin*pd*ph*pw], column order (c,kd,kh,kw); zero padding. */
int clv_patchify(const float* x, void* out_bf16, int B, int Cin, int F, int H, int W, int pd, int ph, int pw, void* stream);
/* Same patch matrix from RAW uint8 frames [B, Cin, F, H, W], normalised on the fly as (v - mean[c]) * inv_std[c]: the
 * GPUNormalize module hook (utils/module_hooks.py:35-87; mean / std of configs/_base_/datasets_local/*.py img_norm_cfg)
 * folded into the load -- 1 byte per sample crosses PCIe instead of 4 and no normalisation pass runs. */
int clv_patchify_u8(const unsigned char* x, const float* mean, const float* inv_std, void* out_bf16, int B, int Cin, int F,
                    int H, int W, int pd, int ph, int pw, void* stream);

/* out[g,c] (+)= scale * sum_{r : (r/div)%mod == g} x[r,c]; out fp32 [mod, C].  nn.Linear bias grads,
 * AdaptiveAvgPool3d (ssl_head.py:105-106), d vis_space_pos / vis_tempor_pos / position embeddings. */
int clv_grouped_colsum(const void* x, int x_is_bf16, long long ld, long long rows, int C, int div, int mod, float scale,
                       float* out, int accumulate, void* stream);

/* y[orow(r),:] = x[irow(r),:] + add0 + bscale * bvec[r / bdiv,:], row(r) = (r/group_rows)*group_stride + r%group_rows + offset
 * (group_rows == 0: identity).  Text half of the fusion concat (cross_transformer.py:84-86,108), its
 * backward slice, and the broadcast of the pooled-feature gradient. */
typedef struct {
  const void* x; int x_is_bf16; long long ld_x;
  long long in_group_rows, in_group_stride, in_offset;
  void* y; int y_is_bf16; long long ld_y;
  long long out_group_rows, out_group_stride, out_offset;
  const float* add0;            /* [C] or NULL */
  const float* bvec; long long bdiv; float bscale;   /* [rows/bdiv, C] or NULL */
  long long rows; int C;
} clv_rows_affine_t;
int clv_rows_affine(const clv_rows_affine_t* desc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stochastic regularisers (training mode).  Counter-based stream: element e keeps iff
 * hash32(seed, e) >= p * 2^32 (splitmix64 finaliser; tests/rng_ref.py restates it in numpy).
 * ------------------------------------------------------------------------------------------- */
/* y[i] = (residual ? residual[i] : 0) + x[i] * keep(offset + i) / (1 - p).   nn.Dropout of HF BertEmbeddings /
 * BertSelfOutput / BertOutput (transformers 4.6.1) and of the heads (ssl_head.py:108-109,211-212,292-293;
 * qa_head.py:12,60).  Backward: the same call on dy with residual == NULL. */
int clv_dropout(const void* x, int x_is_bf16, const void* residual, int residual_is_bf16, void* y, int y_is_bf16,
                long long n, float p, unsigned long long seed, unsigned long long offset, void* stream);
/* out[i] = 1 if element offset + i is kept else 0 (inspection / tests). */
int clv_keep_mask(unsigned char* out, long long n, float p, unsigned long long seed, unsigned long long offset, void* stream);
unsigned int clv_dropout_threshold(float p);
/* host evaluation of the stream (no device needed): the 32 random bits of element idx */
unsigned int clv_rand_u32(unsigned long long seed, unsigned long long idx);
/* y[r,:] = x[r,:] * scale[r / rows_per_group]: per-sample DropPath factor (timm DropPath in
 * swin_transformer_3d.py:499,503) applied to a gradient; the forward factor is clv_gemm_epilogue_t.row_scale. */
int clv_rows_scale(const void* x, int x_is_bf16, void* y, int y_is_bf16, long long rows, int C, const float* scale,
                   long long rows_per_group, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Retrieval evaluation on the device.  Replaces the numpy path of
 * core/evaluation/accuracy.py:427-456 (recall_for_video_text_retrieval): scores = normalize(text) . normalize(video)^T
 * (mmaction/utils/numpy_norm.py:5-8: zero rows are left as they are), then the position of the ground-truth column in
 * the descending order of every row (np.argsort(-scores, axis=1); ties keep column order).  R@k / MedR are integer
 * reductions of rank_out done by the caller.
 * ------------------------------------------------------------------------------------------- */
int clv_cosine_scores(const float* a, int n_a, const float* b, int n_b, int D, float* scores, long long ld_scores,
                      float* workspace /* (n_a + n_b) * D floats */, void* stream);
int clv_retrieval_ranks(const float* scores, long long ld, int rows, int cols, const int* gt_col /* NULL: column == row */,
                        int* rank_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step.  One multi-tensor AdamW pass over the fp32 master weights that also unscales / clips the
 * gradients by their global norm, skips the step on a non-finite norm and refreshes the bf16 operand copies.  Replaces
 * core/hooks/mmcv_Fp16OptimizerHook.py:96-149 + torch.optim.AdamW with paramwise lr / weight decay
 * (configs/exp_local/pretrain_webvid_cc3m.py:129-137).  All tables live in DEVICE memory (the caller uploads them):
 *   tensors_dev[n]               one entry per parameter (param_bf16 may be NULL)
 *   chunk_tensor_dev[n_chunks]   tensor index of every chunk;  chunk_offset_dev[n_chunks] first element of the chunk
 *   status_dev fp32[3]           out: [0] gradient norm (after grad_scale), [1] 1 if the step was skipped, [2] scratch
 *   step_dev fp32[1] or NULL     device-resident count of APPLIED updates (torch.optim.AdamW's state['step']): when given,
 *                                the bias corrections use *step_dev + 1 and the counter advances only if the step was not
 *                                skipped (so `step` is ignored and a resume / an overflow skip keep the schedule exact);
 *                                NULL: the host passes `step` >= 1
 * g' = g * grad_scale * min(1, max_grad_norm / (norm + 1e-6)) (max_grad_norm <= 0: no clipping).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  void* param_bf16;            /* bf16 copy refreshed in the same pass, or NULL */
  long long numel;
  float lr, weight_decay;
} clv_adamw_tensor_t;
int clv_adamw_step(const clv_adamw_tensor_t* tensors_dev, const int* chunk_tensor_dev, const long long* chunk_offset_dev,
                   int n_chunks, int chunk_elems, float beta1, float beta2, float eps, int step, float grad_scale,
                   float max_grad_norm, int check_finite, float* status_dev, float* step_dev, void* stream);

/* dst[index[r],:] += src[r,:]  (fp32 atomics; word-embedding gradient of HF BertEmbeddings). */
int clv_scatter_add_rows(const float* src, const long long* index, float* dst, long long rows, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses (fp32, @force_fp32 in the reference).
 * clv_nce_rank_*: ExclusiveNCEwithRankingLoss.forward (losses/contrastive_loss.py:103-161) for nblk == 3
 * (emb = {video, text, text_mask, text_recon}) and NormSoftmaxLoss.forward (:40-68) for nblk == 1
 * (emb = {video, text}); all matrices [Bg, D] fp32, already all-gathered.  out_losses[0] = nce
 * (loss_v + loss_t), out_losses[1] = pair-wise ranking hinge on the /t-scaled diagonals (:155-159).
 * The workspace (clv_nce_workspace_floats) carries the saved statistics from fwd to bwd.
 * ------------------------------------------------------------------------------------------- */
long long clv_nce_workspace_floats(int nblk, int Bg, int D);
int clv_nce_rank_fwd(const float* const* emb, int nblk, int Bg, int D, float temperature, float margin, int use_rank,
                     float eps, float* workspace, float* out_losses, void* stream);
int clv_nce_rank_bwd(int nblk, int Bg, int D, float temperature, int use_rank, float* workspace, const float* g_nce,
                     const float* g_rank, float* const* grads, void* stream);

/* SoftmaxFocalLossMultiClass.forward (losses/focal_loss.py:61-72; gamma == 0 gives the mean cross-entropy
 * of CrossEntropyLoss._forward, cross_entropy_loss.py:74-81).  Rows whose target == ignore_index are
 * skipped (the row selection of multimodal_transformer_pretrain.py:137-139).  stats fp32 [rows,3],
 * sums fp32 [2] carry state to the backward, which writes d logits (bf16 or fp32, [rows, Vpad]). */
int clv_softmax_focal_fwd(const float* logits, long long ld, long long rows, int V, const long long* target,
                          long long ignore_index, float gamma, float* stats, float* sums, float* loss, void* stream);
int clv_softmax_focal_bwd(const float* logits, long long ld, long long rows, int V, int Vpad, const long long* target,
                          float gamma, const float* stats, const float* sums, const float* g_loss, void* dlogits,
                          int dlogits_is_bf16, long long ld_d, void* stream);

/* Backward of the head_dim-32 window attention on tcgen05 / TMEM (33 <= seq <= 224, no key_mask); same outputs as
 * clv_attention_bwd.  workspace: clv_attention_bwd_tc_workspace_bytes(desc, dbias_table != NULL) bytes. */
long long clv_attention_bwd_tc_workspace_bytes(const clv_attn_desc_t* desc, int with_dbias);
int clv_attention_bwd_tc(const clv_attn_desc_t* desc, const void* qkv, const void* out, const void* dout, const float* lse,
                         void* dqkv, float q_scale, float* dbias_table, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Window attention specialised for the windows Clover actually runs (WindowAttention3D.forward,
 * swin_transformer_3d.py:379-397): head_dim 32, spatial window 7x7, tokens in (d,h,w) order, seq = 49*wd.
 * Same tensors as clv_attention_fwd_tc / clv_attention_bwd_tc (packed qkv, out, lse, dqkv, dbias_table).
 * The relative-position index (:345-359) is evaluated from compile-time token codes; the shift mask
 * (compute_mask :548-562) arrives as two bf16 tables [nwin, seq, 16] that ride in a K-extension of Q K^T:
 *   q_ext row = (0,0,0,0, 10*onehot3(rd), 10*onehot3(rh), 10*onehot3(rw),    1,   1, 0)
 *   k_ext row = (1,1,0,0, 10*onehot3(rd), 10*onehot3(rh), 10*onehot3(rw), -256, -44, 0)
 * with (rd,rh,rw) the per-axis region of the token (region id = 9 rd + 3 rh + rw); tokens whose regions differ in
 * k axes get -100 k, which leaves the softmax exactly like the reference's -100.  NULL tables = no shift mask.
 * forward: wd even in [2, 8]; backward: wd in {2, 4, 8} (wd == 8, the full (8,7,7) window of 16-frame clips and of
 * BASELINE config c2, runs the 196-query kernel on (window, head, query half) units and adds the two dK | dV partials).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int batch, heads, wd;                  /* windows, heads, temporal window extent (seq = 49*wd) */
  const float* bias_table; int table_len; int cfg_wd;   /* fp32 [(2*cfg_wd-1)*169, heads]; configured temporal window */
  const void* q_ext; const void* k_ext; int nwin;       /* bf16 [nwin, seq, 16] each, or NULL; window type = batch % nwin */
} clv_attn_w7_desc_t;

long long clv_attention_w7_fwd_workspace_bytes(const clv_attn_w7_desc_t* desc);
int clv_attention_w7_fwd(const clv_attn_w7_desc_t* desc, const void* qkv, void* out, float* lse,
                         void* workspace /* 16-byte aligned scratch */, void* stream);
long long clv_attention_w7_bwd_workspace_bytes(const clv_attn_w7_desc_t* desc, int with_dbias);
int clv_attention_w7_bwd(const clv_attn_w7_desc_t* desc, const void* qkv, const void* out, const void* dout, const float* lse,
                         void* dqkv, float q_scale, float* dbias_table, void* workspace /* 256-byte aligned */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLOVER_B200_H_ */
