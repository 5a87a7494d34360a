/*
 * x2k.h — C ABI of the B200-native X²-VLM hot path (libx2k.so).
 *
 * The reference (zengyan-97/X2-VLM) has no FFI layer: its hot path is the Python
 * nn.Module surface of models/beit2.py and models/xbert.py executing ATen ops.
 * Each entry point below replaces one ATen op *sequence* of the reference; the
 * file:line it replaces is cited on every function.  Conventions:
 *   - every function returns int: 0 = ok, <0 = error (text via x2k_last_error());
 *   - no C++ exceptions cross the boundary, no torch types in signatures;
 *   - all device buffers are owned by the caller (raw pointers + explicit sizes);
 *   - every launch takes an explicit cudaStream_t (passed as void*);
 *   - functions are re-entrant and hold no mutable global state (the only cached
 *     values are immutable device properties and driver entry points).
 * bf16 tensors are passed as `const void*` / `void*` (2 bytes per element).
 */
#ifndef X2K_H_
#define X2K_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X2K_VERSION 100 /* 0.1.0 */

#define X2K_OK 0
#define X2K_ERR_ARG (-1)
#define X2K_ERR_CUDA (-2)
#define X2K_ERR_UNSUPPORTED (-3)

/* Library version (X2K_VERSION of the build). */
int x2k_version(void);
/* Thread-local text of the last error on this thread ("" if none). */
const char* x2k_last_error(void);
/* Number of kernels launched by this library in this process (monotonic counter, for
 * bench.py's `gpu_launches`). */
int64_t x2k_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Fused GEMM on tcgen05 tensor cores:  C[M,N] = epilogue( A[M,K] · B[N,K]^T )
 *
 * A and B are bf16.  "K-major" (x_mn_major = 0) means the matrix is stored row-major with the
 * contraction dimension contiguous: A as [M,K] (ld = lda), B as [N,K] (ld = ldb) — the
 * F.linear(x, W) layout.  "MN-major" (x_mn_major = 1) means the matrix is stored transposed:
 * A as [K,M] row-major (M contiguous, ld = lda), B as [K,N] row-major (N contiguous).  This
 * covers forward (x·Wᵀ), dgrad (dy·W) and wgrad (dyᵀ·x) of every Linear on the hot path
 * without materialising a transpose.
 *
 * Epilogue, applied per element (m,n) in this order (each step optional):
 *   v  = acc
 *   v += bias[n]
 *   if preact_out:  preact_out[m,n] = bf16(v)            (saved for GELU backward)
 *   if act == X2K_ACT_GELU:       v = gelu_erf(v)
 *   if act == X2K_ACT_GELU_BWD:   v = v * gelu_erf'(aux[m,n])   (aux = saved pre-activation)
 *   if act == X2K_ACT_GELU_SAVE_GRAD:  preact_out[m,n] = bf16(gelu_erf'(v)) INSTEAD of bf16(v), then v = gelu_erf(v)
 *                                 (needs preact_out; the derivative is evaluated once, next to the activation)
 *   if act == X2K_ACT_MUL_AUX:    v = v * aux[m,n]              (aux = derivative saved by GELU_SAVE_GRAD)
 *   if dropout_p > 0:  v = keep(e) ? v*s : 0,  e = m*N+n     (16 random bits per element from
 *                      Philox4x32-7(key = seed, counter = offset + e/8): word (e%8)/2, low half for
 *                      even e; keep iff bits >= thr = round(p*65536); s = 65536/(65536-thr))
 *   if gamma:     v *= gamma[n]                           (BEiT LayerScale)
 *   if row_scale: v *= row_scale[m / rows_per_scale]      (per-sample DropPath keep/(1-p))
 *   if residual:  v += residual[m,n]                      (fp32)
 *   if accumulate (fp32 output only): v += out_f32[m,n]
 *   store to out_bf16 and/or out_f32.
 *
 * Replaces: F.linear / nn.Linear + bias + GELU + LayerScale + DropPath + residual of
 *   models/beit2.py:61-68,127-133,160-161,206-207 and the dense(+dropout+residual) /
 *   dense+GELU of models/xbert.py:339-362,427-431,496-499,511-515 (and their autograd
 *   backward: dgrad, wgrad).
 * ------------------------------------------------------------------------------------------ */
#define X2K_ACT_NONE 0
#define X2K_ACT_GELU 1
#define X2K_ACT_GELU_BWD 2
#define X2K_ACT_GELU_SAVE_GRAD 3
#define X2K_ACT_MUL_AUX 4

typedef struct X2kGemmArgs {
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  int32_t M, N, K;
  int64_t lda, ldb; /* leading dimensions in elements (multiples of 8) */
  int32_t a_mn_major, b_mn_major;
  /* epilogue */
  const float* bias;       /* [N] or NULL */
  int32_t act;             /* X2K_ACT_* */
  const void* aux;         /* bf16 [M, ld_aux], for X2K_ACT_GELU_BWD / X2K_ACT_MUL_AUX */
  int64_t ld_aux;
  void* preact_out;        /* bf16 [M, ld_preact] or NULL */
  int64_t ld_preact;
  float dropout_p;         /* 0 = off */
  uint64_t dropout_seed, dropout_offset;
  const float* gamma;      /* [N] or NULL */
  const float* row_scale;  /* [ceil(M / rows_per_scale)] or NULL */
  int32_t rows_per_scale;
  const float* residual;   /* fp32 [M, ld_res] or NULL */
  int64_t ld_res;
  int32_t accumulate;      /* out_f32 += result */
  void* out_bf16;          /* bf16 [M, ld_out_bf16] or NULL */
  int64_t ld_out_bf16;
  float* out_f32;          /* fp32 [M, ld_out_f32] or NULL */
  int64_t ld_out_f32;
  int32_t tile_n;          /* 0 = auto, else 128 or 256 */
  int32_t max_ctas;        /* 0 = all SMs */
  int32_t split_k;         /* 0 = auto (split K over CTAs for wgrad-shaped GEMMs with a plain fp32 output,
                              accumulated with atomics), 1 = never, n = force n slices */
  const uint64_t* dropout_offset_dev; /* optional device counter ADDED to dropout_offset at run time: a captured
                              CUDA graph advances it between replays so every step draws fresh masks */
  /* Fused cross entropy over N (the vocabulary GEMM of the MLM / LM heads, models/xbert.py:805-834 + :1653-1661):
   *   ce_mode 1: no logits are stored; after the bias add every 16-column group of row m writes its online-softmax
   *              statistics (max, sum exp(v - max)) to ce_partials[m, n/16, 0..1] and the label's logit to
   *              ce_target_logit[m]; x2k_ce_finalize turns them into the row's log-sum-exp and loss;
   *   ce_mode 2: (backward) the logits are recomputed and out_bf16[m,n] = ce_row_grad[m] * (exp(v - ce_lse[m]) - [n == label]).
   * ce_labels: int64 [M], negative = ignored row (loss 0, no target logit). */
  int32_t ce_mode;
  const int64_t* ce_labels;
  float* ce_partials;
  float* ce_target_logit;
  const float* ce_lse;
  const float* ce_row_grad;
} X2kGemmArgs;

int x2k_gemm(const X2kGemmArgs* args, void* stream);

/* Second half of the fused cross entropy: per row m reduce the ceil(N/16) partial statistics of x2k_gemm(ce_mode 1) into
 * lse[m] = log sum_n exp(logit[m,n]) and loss[m] = labels[m] >= 0 ? lse[m] - target_logit[m] : 0 (CrossEntropyLoss with
 * ignore_index, reduction 'none').  Streaming, one warp per row. */
int x2k_ce_finalize(const float* partials, const float* target_logit, const int64_t* labels, int32_t M, int32_t N,
                    float* lse, float* loss, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound): LayerNorm forward / backward.
 *
 * x2k_layernorm_fwd: y = (x - mean) * rstd * w + b over the last dim D of x[M,D] (fp32 in),
 *   writes any of y_bf16 / y_f32, and mean/rstd [M] for backward.
 *   Replaces nn.LayerNorm at models/beit2.py:193,201,207,411 (eps 1e-6) and
 *   models/xbert.py:212,430,514 (eps 1e-12).
 * x2k_layernorm_bwd: dx = LN'(dy) (+ dx_residual if given); dw/db are ACCUMULATED (+=) into
 *   fp32 [D] buffers.  dy = dy_f32 + dy_bf16 (either may be NULL, not both) — the fp32 term is
 *   the residual-stream gradient, the bf16 term a branch gradient produced by a dgrad GEMM.
 * ------------------------------------------------------------------------------------------ */
int x2k_layernorm_fwd(const float* x, const float* w, const float* b, int32_t M, int32_t D, float eps,
                      void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
int x2k_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* w,
                      const float* mean, const float* rstd, const float* dx_residual, int32_t M,
                      int32_t D, float* dx, float* dw, float* db, void* stream);
/* x2k_layernorm_bwd_dropcast: x2k_layernorm_bwd followed, in the same pass, by what x2k_scale_cast_colsum(dx, dropout)
 * computes: g_bf16[m,n] = bf16(dx[m,n] * keepscale(m,n)) (Philox element index m*D+n, seed/offset of the forward
 * GEMM's dropout) and dbias[n] += sum_m dx[m,n] * keepscale(m,n) (dbias may be NULL).  Backward of the post-LN BERT
 * "LayerNorm(dropout(dense(h)) + residual)" (models/xbert.py:427-431, :511-515) without re-reading dx. */
int x2k_layernorm_bwd_dropcast(const void* dy_bf16, const float* dy_f32, const float* x, const float* w,
                               const float* mean, const float* rstd, const float* dx_residual, int32_t M,
                               int32_t D, float* dx, float* dw, float* db, float dropout_p,
                               uint64_t dropout_seed, uint64_t dropout_offset,
                               const uint64_t* dropout_offset_dev, void* g_bf16, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------
 * x2k_scale_cast_colsum: backward of the "dropout / LayerScale / DropPath + residual" epilogue.
 *   g[m,n]   = bf16( dx[m,n] * keepscale(m,n) * gamma[n] * row_scale[m / rows_per_scale] )
 *   dbias[n]  += sum_m g[m,n]                       (if dbias)
 *   dgamma[n] += sum_m dx[m,n] * row_scale[..] * y[m,n]   (if dgamma; y = saved bf16 branch output)
 * keepscale regenerates the Philox mask of x2k_gemm's dropout (same seed/offset, N = ld of the
 * forward GEMM's logical N).
 * Replaces autograd of models/beit2.py:206-207 and models/xbert.py:428-431,512-515.
 * ------------------------------------------------------------------------------------------ */
int x2k_scale_cast_colsum(const float* dx, int64_t ld_dx, int32_t M, int32_t N, const float* gamma,
                          const float* row_scale, int32_t rows_per_scale, float dropout_p,
                          uint64_t dropout_seed, uint64_t dropout_offset,
                          const uint64_t* dropout_offset_dev, const void* y_bf16,
                          int64_t ld_y, void* g_bf16, int64_t ld_g, float* dbias, float* dgamma,
                          void* stream);

/* dcol[n] += sum_m x[m,n] for a bf16 matrix (bias gradients of Linear layers). */
int x2k_colsum_bf16(const void* x_bf16, int64_t ld, int32_t M, int32_t N, float* dcol, void* stream);

/* out[s,:] = sum_{b: index[b]==s} in[b,:] (bf16 in/out, fp32 accumulate; row_elems % 8 == 0).
 * Reduces the per-query-sequence dK/dV of cross-attention onto the shared image K/V
 * (the K/V cache across the ITM/MLM/bbox passes of models/xvlm.py:859-925). */
int x2k_segment_sum_bf16(const void* in_bf16, const int32_t* index, int32_t n_rows, int64_t row_elems,
                         int32_t n_seg, void* out_bf16, void* stream);

/* dst_bf16[i] = bf16(src_f32[i]), i < n (weight shadow copies, activation casts). */
int x2k_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention on tcgen05:  O = softmax(scale·Q·Kᵀ + bias[h] + mask[b]) · V   per (batch b, head h).
 *
 * Q rows of sequence b live at q + (b*Lq + i)*ld_q + h*64 (bf16, head_dim 64 only); K/V rows of
 * sequence b at k + ((kv_index ? kv_index[b] : b)*Lk + j)*ld_k + h*64 — i.e. the kernels read
 * the packed QKV projection output in place ([B·L, 3D] for BEiT/BERT self-attention, separate
 * Q and KV buffers for cross-attention) and several query sequences may share one K/V sequence
 * (the image K/V cache across the ITM/MLM passes, models/xvlm.py:859-899).
 * Any Lq / Lk: up to 256 keys the whole key range of a (sequence, head) stays in TMEM (attn.cu); longer sequences
 * (384 px / 768 px fine-tuning: N = 577 / 2305) run key-blocked kernels with an online softmax (attn_long.cu).
 * Let Lk_pad = Lk rounded up to 16 and Lk_32 = Lk rounded up to 32.
 * bias: fp32, element (h,i,j) at bias[h*bias_h_stride + i*bias_q_stride + j] (BEiT relative
 *   position bias already gathered), or NULL.  Strides are multiples of 4, q stride >= Lk_32
 *   (the kernels stage it in coalesced 32-row x 32-column blocks).
 * mask: fp32 additive, element (b,i,j) at mask[b*mask_b_stride + i*mask_q_stride + j]
 *   (mask_q_stride = 0 for a key mask [B,Lk]), or NULL.  Same stride rules.
 * dropout on probabilities (BERT attention_probs_dropout): element e = ((b*H+h)*Lq + i)*Lk_pad + j
 *   of the Philox stream described at x2k_gemm.
 * Outputs: O bf16 at o + (b*Lq+i)*ld_o + h*64, lse fp32 [B,H,Lq] (log2 domain: log2 sum exp2).
 * Replaces models/beit2.py:135-159 and models/xbert.py:364-410.
 *
 * x2k_attn_bwd: given dO (and the forward's o, lse), recomputes P and produces dQ, dK, dV (bf16,
 * addressed like q/k/v through their own pointers/ld; dK/dV are per *query batch* b — the caller
 * reduces over shared K/V).  If ds_out != NULL also exports dS (bf16) at
 * ds_out[b*ds_b_stride + h*ds_h_stride + i*ds_q_stride + j] for the bias gradient.
 * ------------------------------------------------------------------------------------------ */
typedef struct X2kAttnArgs {
  const void *q, *k, *v; /* bf16 */
  int64_t ld_q, ld_k, ld_v;
  int32_t B, H, Lq, Lk;
  const int32_t* kv_index; /* [B] or NULL */
  int32_t n_kv;            /* number of K/V sequences behind k/v (0 = B) */
  float scale;
  const float* bias;
  int64_t bias_h_stride, bias_q_stride;
  const float* mask;
  int64_t mask_b_stride, mask_q_stride;
  float dropout_p;
  uint64_t dropout_seed, dropout_offset;
  void* o;                 /* bf16 (output of fwd, input of bwd) */
  int64_t ld_o;
  float* lse;              /* [B,H,Lq] (output of fwd, input of bwd) */
  /* backward only */
  const void* d_o;         /* bf16 */
  int64_t ld_do;
  void *dq, *dk, *dv;      /* bf16 */
  int64_t ld_dq, ld_dk, ld_dv;
  void* ds_out;            /* bf16 or NULL */
  int64_t ds_b_stride, ds_h_stride, ds_q_stride;
  /* grouped cross-attention (optional): work-item table built by x2k_attn_group_build from kv_index.
   * When set, x2k_attn_bwd writes dK/dV PER K/V SOURCE: n_kv*Lk rows addressed like k/v, already
   * summed over the query sequences that share the source (zeros for a source nobody reads). */
  const int32_t* kv_groups;
  const uint64_t* dropout_offset_dev; /* optional device counter added to dropout_offset (see X2kGemmArgs) */
  /* backward, optional workspace: fp32 [B,H,Lq].  When given, x2k_attn_bwd first runs a streaming pre-kernel that
   * writes delta[b,h,i] = sum_d O[b,i,h,d] * dO[b,i,h,d] there, and the attention kernels read one float per row
   * instead of fetching the row's O and dO at their start (a DRAM round trip on every CTA's critical path). */
  float* delta_ws;
  /* backward, workspace of the key-blocked kernels (Lq > 256 or Lk > 256): x2k_attn_bwd_workspace_bytes(args) bytes,
   * 16-byte aligned; fp32 [B*Lq, H*64] into which the key-block CTAs reduce their dQ contributions.  The call zeroes
   * it on the stream.  NULL is fine whenever the query says 0 bytes. */
  float* dq_ws;
} X2kAttnArgs;

int x2k_attn_fwd(const X2kAttnArgs* args, void* stream);
int x2k_attn_bwd(const X2kAttnArgs* args, void* stream);
/* Bytes of X2kAttnArgs.dq_ws that x2k_attn_bwd needs for these shapes (B, H, Lq, Lk of *args; 0 = none). */
int64_t x2k_attn_bwd_workspace_bytes(const X2kAttnArgs* args);

/* Attention maps on request (the fused kernels never materialise them):
 *   mode 0: out[b,h,i,j] = softmax_j(scale·q_i·k_j + bias + mask)  — the PRE-dropout probabilities the reference returns as
 *           `attn_prob` / `attention_probs` with output_attentions (models/beit2.py:152,166; models/xbert.py:392,410) and
 *           stores with save_attention (models/xbert.py:394-396); needs q, k, lse (from x2k_attn_fwd), bias / mask;
 *   mode 1: out[b,h,i,j] = (dO_i·v_j) * keepscale(b,h,i,j) — dL/d(attention_probs), what the reference's
 *           register_hook(save_attn_gradients) receives; needs d_o, v and the forward's dropout parameters.
 * out: fp32 [B, H, Lq, Lk] contiguous.  HBM-bound on the output. */
int x2k_attn_probs(const X2kAttnArgs* args, int32_t mode, float* out, void* stream);

/* Short query sequences (Lq rounded up to 8 <= 64, no bias): several sequences share one 128-row MMA
 * tile.  Self-attention (kv_index == NULL, Lq == Lk-sized segments) is packed automatically inside
 * x2k_attn_fwd/bwd.  Cross-attention over shared K/V needs the sequences grouped by K/V source:
 *   x2k_attn_group_slots(Lq)            -> sequences per tile G (0 = Lq not eligible)
 *   x2k_attn_group_table_ints(B,n_kv,Lq)-> int32 elements the caller allocates for the table
 *   x2k_attn_group_build(...)           -> fills the table on the device (one launch, no host sync):
 *       {n_items, G, n_kv, B} | first[n_kv+1] | pad to 4 | items[(B + n_kv*(G-1))/G][12] = {src, n, b[0..7], 0, 0}
 *       sequences of a source keep their batch order (deterministic).
 * Replaces the per-(text, image) repeated K/V handling of models/xvlm.py:859-899 + models/xbert.py:343-349. */
int32_t x2k_attn_group_slots(int32_t Lq);
int64_t x2k_attn_group_table_ints(int32_t B, int32_t n_kv, int32_t Lq);
int x2k_attn_group_build(const int32_t* kv_index, int32_t B, int32_t n_kv, int32_t Lq, int32_t* table,
                         void* stream);

/* BEiT relative position bias (models/beit2.py:138-143): out[h, i, j] = table[index[i*N+j], h]
 * written with row stride ld_out (>= N, multiple of 4), h stride N*ld_out; and its backward:
 * dtable[index[i*N+j], h] += sum_b ds[b,h,i,j]  (ds = bf16 dS exported by x2k_attn_bwd). */
int x2k_relpos_bias_gather(const float* table, const int64_t* index, int32_t N, int32_t H,
                           float* out, int64_t ld_out, void* stream);
int x2k_relpos_bias_scatter(const void* ds_bf16, int32_t B, int32_t H, int32_t N, int64_t ds_b_stride,
                            int64_t ds_h_stride, int64_t ds_q_stride, const int64_t* index,
                            float* dtable, void* stream);

/* ------------------------------------------------------------------------------------------
 * The two small HBM-bound ends of the encoders.  D must be a multiple of 128, <= 1024.
 *
 * x2k_embed_ln_fwd: y[m] = dropout(LayerNorm(word[ids[m]] + pos[p(m)] + type[t(m)])), t(m) = type_ids ? type_ids[m] : 0,
 *   p(m) = pos_ids ? pos_ids[m] : pos_offset + m % L.  Writes y_f32 (and y_bf16 if given) and the LayerNorm statistics
 *   mean / rstd [M] for backward.  Dropout element index m*D + n of the Philox stream described at x2k_gemm.
 *   Replaces BertEmbeddings.forward (models/xbert.py:189-216): three gathers, two adds, LayerNorm, dropout.
 * x2k_embed_ln_bwd: recomputes the summed row, applies dropout' and LayerNorm' and ACCUMULATES (+=, vector reductions)
 *   the row gradient into dword[ids[m]], dpos[p(m)], dtype[t(m)]; dw / db (LayerNorm weight / bias) are accumulated too.
 * ------------------------------------------------------------------------------------------ */
int x2k_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int32_t M, int32_t D,
                     int32_t L, int32_t pos_offset, const float* word, const float* pos, const float* type,
                     const float* ln_w, const float* ln_b, float eps, float dropout_p, uint64_t dropout_seed,
                     uint64_t dropout_offset, const uint64_t* dropout_offset_dev, float* y_f32, void* y_bf16,
                     float* mean, float* rstd, void* stream);
int x2k_embed_ln_bwd(const float* dy, const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids, int32_t M,
                     int32_t D, int32_t L, int32_t pos_offset, const float* word, const float* pos, const float* type,
                     const float* ln_w, const float* ln_b, const float* mean, const float* rstd, float dropout_p,
                     uint64_t dropout_seed, uint64_t dropout_offset, const uint64_t* dropout_offset_dev, float* dword,
                     float* dpos, float* dtype, float* dw, float* db, void* stream);

/* x2k_pool_tail_fwd: the tail of VisionTransformer.forward (models/beit2.py:409-436).  x [n_img, N, D] is the output of
 *   the last block (token 0 = cls, dropped).  For every output sequence s (image g = group ? group[s] : s):
 *     out[s, r] = fc_norm(x[g, r]) for r = 1..N-1,   out[s, 0] = sum_r a[s,r] * out[s, r] / sum_r a[s,r]
 *   with a = atts ? atts[s, r] (int64 0/1 region masks, column 0 ignored) : 1 — i.e. the plain mean of the normalised
 *   patch tokens (full-image embeddings) or the mask-weighted mean of region mode (idx_to_group_img + image_atts[:, 1:]).
 *   mean / rstd [n_out, N] (optional, both or none) keep the LayerNorm statistics for backward.
 * x2k_pool_tail_bwd: dx[g, r] += LN'(d_out[s, r] + a[s,r] / sum a * d_out[s, 0]) by vector reductions (dx is zeroed first
 *   unless accumulate != 0), dw / db of fc_norm accumulated. */
int x2k_pool_tail_fwd(const float* x, int32_t n_img, int32_t n_out, int32_t N, int32_t D, const int64_t* group,
                      const int64_t* atts, const float* w, const float* b, float eps, float* out, float* mean,
                      float* rstd, void* stream);
int x2k_pool_tail_bwd(const float* d_out, const float* x, int32_t n_img, int32_t n_out, int32_t N, int32_t D,
                      const int64_t* group, const int64_t* atts, const float* w, const float* mean,
                      const float* rstd, float* dx, int32_t accumulate, float* dw, float* db, void* stream);

/* ------------------------------------------------------------------------------------------
 * Flat-buffer optimizer step (replaces optim.py:26-104 AdamW +
 * accelerators/apex_ddp_accelerator.py:99-102 clip_grad_norm_):
 * x2k_sumsq: out[0] += sum g[i]^2 (fp32, for the global grad norm).
 * x2k_adamw_flat: decoupled-weight-decay Adam over a flat fp32 parameter buffer; per-element
 *   hyper-parameters come from a segment table (seg_end[s] exclusive end offsets, lr[s], wd[s]);
 *   grads are pre-scaled by *grad_scale_dev (clip coefficient computed on device);
 *   also writes the bf16 shadow copy used by the GEMMs.  The step count (bias correction) is
 *   `step`, or *step_dev when step_dev != NULL (so a captured CUDA graph can advance it).
 *   zero_grad != 0: every gradient element is set to 0 right after it has been consumed (the next step's
 *   optimizer.zero_grad() for 4 more bytes per parameter instead of a separate 1 GB fill).
 * ------------------------------------------------------------------------------------------ */
int x2k_sumsq(const float* g, int64_t n, float* out, void* stream);
int x2k_adamw_flat(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n,
                   const int64_t* seg_end, const float* seg_lr, const float* seg_wd, int32_t n_seg,
                   float beta1, float beta2, float eps, int32_t step, const int32_t* step_dev,
                   const float* grad_scale_dev, int32_t zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* X2K_H_ */
