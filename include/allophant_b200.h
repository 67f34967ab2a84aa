/* allophant_b200 — C ABI of the B200-native acoustic-model forward/loss path.
 *
 * This header is the drop-in boundary.  The reference (kgnlp/allophant) has no
 * FFI under `allophant.network`: its hot path is a chain of torch / transformers
 * library calls made from Python.  Each entry point below replaces one such
 * call site (cited per function as `reference file:line`, paths relative to the
 * reference checkout; `HF:` = transformers/models/wav2vec2/modeling_wav2vec2.py).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    parameter name ends in `_host`;
 *  - buffers are caller-owned (the Python host allocates them with torch);
 *  - `stream` is a `cudaStream_t` passed as `void*`; launches are asynchronous;
 *  - return value: APH_OK (0) or a negative APH_ERR_* code; `aph_last_error()`
 *    returns a thread-local human-readable description of the last failure;
 *  - no exceptions cross the boundary, no CPU fallback exists behind it.
 */
#ifndef ALLOPHANT_B200_H_
#define ALLOPHANT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APH_OK 0
#define APH_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer) */
#define APH_ERR_CUDA (-2)        /* CUDA runtime / driver error                    */
#define APH_ERR_UNSUPPORTED (-3) /* valid request outside the implemented envelope */

#define APH_ABI_VERSION 7

/* ---- library ------------------------------------------------------------ */
int aph_abi_version(void);
const char* aph_last_error(void);
/* Number of kernels launched by this library since the last reset (bench.py
 * reports it as `gpu_launches`). */
int64_t aph_launch_count(void);
void aph_reset_launch_count(void);
/* Programmatic dependent launch between the library's kernels (each kernel's prologue overlaps the tail of its
 * predecessor in the stream; results are identical).  On by default, APH_PDL=0 in the environment or this call turn it
 * off.  Returns the previous setting.  No reference counterpart: torch launches its kernels in plain stream order. */
int aph_set_pdl(int enabled);
/* Tail split of aph_gemm_bf16: when the last wave of 256 x 256 output tiles fills at most half of the cluster slots, each of
 * its tiles is computed as two 256 x 128 halves on two clusters (results bit-identical to the unsplit kernel: every output
 * element accumulates the same products in the same order).  On by default, APH_GEMM_TAIL_SPLIT=0 or this call turn it
 * off; returns the previous setting.  No reference counterpart. */
int aph_set_gemm_tail_split(int enabled);
/* Which kernel aph_attention_bf16* launches: 0 = by problem size (the default: 64-key blocks with two CTAs per SM while a
 * persistent CTA would get at most one (utterance, head, query-tile pair) item, the persistent query-tile-pair kernel beyond),
 * 1 = always the 64-key kernel, 2 = always the pair kernel (APH_ATT_V1=1 / APH_ATT_V1=0 in the environment).  Returns the
 * previous setting; results agree within the tolerances of the tests.  No reference counterpart. */
int aph_set_attention_kernel(int mode);

/* ---- tensor-core GEMM (tcgen05 + TMEM accumulators, TMA-fed) ------------- */
/* One kernel serves every dense contraction of the path:
 *   nn.Linear call sites      HF:429-434 (feature projection), HF:510-515 (q/k/v),
 *                             HF:546 (out_proj), HF:565-572 (feed forward),
 *                             acoustic_model.py:415,298 (classifier heads),
 *                             acoustic_model.py:234 (composition logits)
 *   strided Conv1d layers 1-6 HF:281-299 (implicit GEMM, overlapping-row TMA view)
 *   grouped positional conv   HF:326-368 (sliding-tap implicit GEMM)
 * D[m, n] = sum_k A[m, k] * B[n, k]; A, B bf16; fp32 accumulation.
 * Epilogue (all optional, applied in this order):
 *   v = acc * scale + bias[n]; v = gelu_erf(v); v += resid[m, n];
 *   v = 0 where the row is a padded frame; store fp32 and/or bf16.
 */
#define APH_GEMM_ROWS 0    /* A rows addressed as base + row*a_row_stride (plain & strided conv) */
#define APH_GEMM_TAPS 1    /* sliding taps: k-block j reads rows t - pad + j of channel group n/64 */
#define APH_GEMM_DIAG_TAPS 2 /* weight gradient of the grouped positional conv (both operands MN-major):
                              * work item (tap j, 256-channel block q): D_j[q] = A[:, q]^T . shift_j(B[:, q]) */
#define APH_EPI_STORE 0    /* generic epilogue described above                                   */
#define APH_EPI_QKV 1      /* scatter bf16 into Q[b,h,t,64] (pre-scaled), K[b,h,t,64], V[b,h,t,64] (= args.vmat) */

typedef struct aph_gemm_args {
  /* A operand, bf16. Logical view [batch][a_rows][a_inner]. */
  const void* a;
  int64_t a_row_stride;   /* elements between consecutive rows (may be < a_inner: overlapping rows) */
  int64_t a_batch_stride; /* elements between batches */
  int32_t a_rows;         /* rows per batch (output positions) */
  int32_t a_inner;        /* addressable elements per row */
  int32_t batch;
  int32_t mode;           /* APH_GEMM_ROWS | APH_GEMM_TAPS */
  int32_t tap_pad;        /* TAPS: left padding (64 for the 128-tap positional conv) */
  /* B operand, bf16 [n][k] row-major (nn.Linear weight layout). */
  const void* b;
  int32_t n;
  int32_t k;              /* multiple of 64 */
  /* epilogue */
  int32_t epilogue;       /* APH_EPI_* */
  int32_t gelu;           /* activation after bias: 0 none, 1 GELU (erf), 2 ReLU, 3 LeakyReLU(0.01) */
  float scale;
  const float* bias;      /* [n] or NULL */
  const float* resid;     /* fp32 [rows][ld_resid] or NULL; may alias out_f32 */
  int64_t ld_resid;
  float* out_f32;         /* or NULL */
  int64_t ld_f32;
  void* out_bf16;         /* or NULL */
  int64_t ld_bf16;
  int64_t out_batch_rows; /* output row = b * out_batch_rows + t */
  const int32_t* lengths; /* valid frames per utterance or NULL */
  int32_t len_period;     /* rows per utterance: utt = row / len_period, t = row % len_period */
  /* APH_EPI_QKV */
  void* q;
  void* kmat;
  void* vt;               /* optional: V transposed [b,h,64,t_v] as well (NULL = not written; nothing in the library reads it) */
  int32_t heads;
  int32_t t_v;            /* padded key length of vt (multiple of 8) */
  float q_scale;
  /* ---- ABI 2: operand majors and training epilogues (all zero = ABI 1 behaviour) ---------------
   * MN-major operands let the backward GEMMs of nn.Linear read activations and weights in the
   * layout the forward pass left them in (no transposed copies):
   *   dgrad  dX[m,i] = sum_o dY[m,o] W[o,i]   A = dY (K-major),  B = W  [k=o][n=i]  (b_mn_major)
   *   wgrad  dW[o,i] = sum_m dY[m,o] X[m,i]   A = dY [k=m][m=o]  (a_mn_major), B = X [k=m][n=i]
   * An MN-major operand is addressed as [k_batch][k_seq][cols]: element (segment s, row r, col c) at
   * base + s*seg_stride + r*row_stride + c; the contraction runs over all (s, r); rows outside
   * [0, k_seq) read as zero (TMA fill).  k = k_batch * k_seq is implied, args.k is ignored. */
  int32_t a_mn_major;     /* A stored [k rows][a_rows cols], row stride a_row_stride, segment stride a_batch_stride */
  int32_t b_mn_major;     /* B stored [k rows][n cols] */
  int64_t b_row_stride;   /* elements between stored rows of B (0: k for K-major, n for MN-major) */
  int64_t b_seg_stride;   /* MN-major B: elements between segments */
  int32_t k_seq;          /* MN-major: stored rows per segment */
  int32_t k_batch;        /* MN-major: number of segments (0 = 1) */
  int32_t b_k_shift;      /* MN-major B: row offset added inside the segment (DIAG_TAPS: tap - tap_pad is added on top) */
  int32_t n_taps;         /* DIAG_TAPS: number of taps; output [n_taps][a_rows][256] fp32 */
  void* aux_bf16;         /* store epilogue: value BEFORE gelu (after scale/bias), bf16 [rows][ld_aux] or NULL */
  int64_t ld_aux;
  const void* gelu_bwd;   /* store epilogue: v *= act'(pre[row][col]) with pre bf16 [rows][ld_gelu_bwd] or NULL; the
                           * activation is GELU unless act_bwd (last field) says ReLU (2) / LeakyReLU (3) */
  int64_t ld_gelu_bwd;
  void* vmat;             /* APH_EPI_QKV: V row-major [b,h,t,64] (required) */
  /* train-mode dropout of (acc*scale + bias [gelu]) BEFORE the residual is added (HF hidden_dropout of the attention
   * and feed-forward blocks): keep iff the element's 16-bit hash >= drop_threshold (0 = off), kept values * drop_scale.
   * The mask is a function of (drop_seed, output row, output column) only: see aph_dropout_2d. */
  uint32_t drop_threshold;
  uint32_t drop_seed;
  float drop_scale;
  int32_t act_bwd;        /* activation whose derivative gelu_bwd applies: 0 / 1 GELU (erf), 2 ReLU, 3 LeakyReLU(0.01) */
  /* ---- ABI 5: LayerNorm folded into the GEMMs around it (inference; all zero = off) -------------------------------
   * The pre-LN encoder computes y = Linear(LayerNorm(x)) with x the fp32 residual stream the previous Linear just wrote
   * (HF:730-756: layer_norm -> attention, final_layer_norm -> feed_forward).  Instead of a LayerNorm kernel between them,
   *   producer  (fp32 output + residual): also stores a bf16 copy of the output (out_bf16 / ld_bf16, through TMA) and, per
   *             output row and column slot s = column / 128, the pair (sum, sum of squares) of the stored values in
   *             row_stats[row][s] (float2); row_stats_slots = 2 * ceil(n / 256);
   *   consumer  reads that bf16 copy as A, a weight with gamma folded in (B[n][k] = W[n][k] * gamma[k]), bias' = bias +
   *             W beta, and applies y = rstd * (acc - mean * ln_colsum[n]) + bias'[n] with mean / rstd from ln_stats
   *             (summed over ln_slots) over ln_cols columns; ln_colsum[n] = sum_k B[n][k] (aph_fold_layernorm_linear). */
  float* row_stats;
  int32_t row_stats_slots;
  const float* ln_stats;
  int32_t ln_slots;
  int32_t ln_cols;
  const float* ln_colsum;
  float ln_eps;
  /* APH_GEMM_TAPS with groups that are not 64 channels wide (wav2vec2-base: 768 channels in 16 groups of 48, HF:326-368): the
   * grouped conv is run block-diagonally over super groups of taps_span channels (a common multiple of the group width and 64:
   * 192), k = taps * taps_span, and B is packed [n][tap][taps_span] with zeros where input and output channel are in different
   * groups.  0 = 64 (one 64-channel group per output tile). */
  int32_t taps_span;
} aph_gemm_args;

int aph_gemm_bf16(const aph_gemm_args* args, void* stream);

/* ---- variable-length self-attention (tcgen05, flash-style online softmax) --- */
/* Replaces the SDPA call of Wav2Vec2Attention (HF:466-549) and the dense
 * additive mask of HF:758-762: keys t >= lengths[b] get probability 0.
 *   q, k, v : bf16 [n_utt*heads][T][64] (q already scaled by head_dim^-0.5 * log2(e): the kernel exponentiates
 *          with exp2; v row-major, read as an MN-major tensor-core operand: no transposed copy)
 *   ctx  : bf16 [n_utt*T][heads*64]    rows of query tiles that hold padded frames only are set to zero
 */
int aph_attention_bf16(const void* q, const void* k, const void* v, void* ctx,
                       const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T, void* stream);
/* Training variant: additionally writes lse2[n_utt*heads][T] = log2-domain log-sum-exp of every
 * query row of a non-skipped tile (NULL = same as aph_attention_bf16). */
int aph_attention_bf16_lse(const void* q, const void* k, const void* v, void* ctx, float* lse2,
                           const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T, void* stream);
/* Backward of the same call (autograd of HF:466-549, reached from loss.backward(), estimator.py:738).
 *   q, k, v : bf16 [n_utt*heads][T][64] as written by the QKV epilogue (q pre-scaled)
 *   ctx, d_ctx : bf16 [n_utt*T][heads*64] forward output and its gradient (rows of padded frames must be 0 in d_ctx)
 *   lse2 : from aph_attention_bf16_lse; delta_scratch : fp32 [n_utt*heads][T]
 *   dqkv : bf16 [n_utt*T][3*heads*64] = (dQ | dK | dV), dQ w.r.t. the UNSCALED query projection;
 *          rows of padded frames are written as zeros. */
int aph_attention_backward_bf16(const void* q, const void* k, const void* v, const void* ctx,
                                const void* d_ctx, const float* lse2, float* delta_scratch,
                                void* dqkv, const int32_t* lengths, int32_t n_utt, int32_t heads,
                                int32_t T, void* stream);

/* Train-mode variants with attention dropout (HF Wav2Vec2Attention, `attention_dropout`): the softmax probabilities
 * are dropped with a counter-based keep mask — element (b*heads+h, q, k) is kept iff the 16-bit half (k & 1) of
 * hash(seed, (b*heads+h)*T + q, k >> 1) is >= drop_threshold (= round(p * 65536); 0 = no dropout) — and kept values are
 * scaled by drop_scale = 1/(1-p).  The backward pass regenerates the mask from the same (threshold, seed). */
int aph_attention_bf16_dropout(const void* q, const void* k, const void* v, void* ctx, float* lse2,
                               const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                               uint32_t drop_threshold, uint32_t drop_seed, float drop_scale, void* stream);
int aph_attention_backward_bf16_dropout(const void* q, const void* k, const void* v, const void* ctx,
                                        const void* d_ctx, const float* lse2, float* delta_scratch, void* dqkv,
                                        const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                                        uint32_t drop_threshold, uint32_t drop_seed, float drop_scale,
                                        void* stream);

/* ---- train-mode stochastic regularisation (HF modeling_wav2vec2.py dropout sites, SpecAugment; -------------
 * ---- acoustic_model.py:486-488) -------------------------------------------------------------------------- */
/* out = x * keep * scale for fp32 x [rows][ld_x] (cols % 4 == 0): element (row, col) is kept iff the 16-bit half
 * (col & 1) of hash(seed, row, col >> 1) >= threshold.  Rows flagged in row_mask (uint8 [rows], optional: SpecAugment)
 * are replaced by row_fill[cols] (NULL: zeros) instead.  out_f32 may alias x; out_bf16 optional.  The same call applied
 * to a gradient is the backward pass. */
int aph_dropout_2d(const float* x, int64_t ld_x, int64_t rows, int32_t cols, uint32_t threshold, uint32_t seed,
                   float scale, const uint8_t* row_mask, const float* row_fill, float* out_f32, int64_t ld_f32,
                   void* out_bf16, int64_t ld_bf16, void* stream);
/* bf16 in place (blocks of the classifier feature matrix, acoustic_model.py:486-488) */
int aph_dropout_bf16_2d(void* x_bf16, int64_t ld, int64_t rows, int32_t cols, uint32_t threshold, uint32_t seed,
                        float scale, void* stream);
/* SpecAugment time mask (HF _compute_mask_indices): uint8 mask [n_utt][seq], spans of mask_length frames drawn
 * without replacement inside each utterance's frames[i] valid frames. */
int aph_spec_augment_mask(const int32_t* frames, int32_t n_utt, int32_t seq, float mask_prob, int32_t mask_length,
                          int32_t min_masks, uint32_t seed, uint8_t* mask, void* stream);
/* SpecAugment along the feature axis (HF mask_feature_prob): x[n][t][c] = 0 where col_mask[n][c] (uint8 [n_utt][cols], made
 * by aph_spec_augment_mask with seq = cols); fp32 in place + optional bf16 copy.  The same call on a gradient is the backward. */
int aph_mask_columns(float* x, int64_t ld, int32_t n_utt, int32_t seq, int32_t cols, const uint8_t* col_mask,
                     void* x_bf16, int64_t ld_bf16, void* stream);
/* backward of the row replacement: d_fill[col] = sum of d over flagged rows; flagged rows of d are zeroed */
int aph_masked_rows_backward(float* d, int64_t ld, int64_t rows, int32_t cols, const uint8_t* row_mask,
                             float* d_fill, void* stream);


/* Debug aid: progress markers of block (0,0) of the attention backward kernels are written to this
 * host-mapped int32[16] array (NULL = off, the default). */
int aph_debug_set_progress(int32_t* host_mapped);
/* Debug aid: clock64 stamps of CTA (0,0) of the attention forward kernel go to this DEVICE int64[32] array (NULL = off). */
int aph_debug_set_timeline(int64_t* device_buffer);

/* ---- waveform normalisation and frame bookkeeping ------------------------- */
/* zero_mean_unit_var_norm, acoustic_model.py:762-767:
 *   mean = sum_all(x) / len; var = sum_valid((x-mean)^2) / len;
 *   out  = (x - mean) / sqrt(var + 1e-7) on valid samples, 0 on padding.
 * aph_wave_stats writes mean_rstd[n_utt][2] (fp32) using stats_scratch[n_utt*3] (fp64);
 * aph_wave_norm materialises the normalised waveform (the fused encoder never does:
 * aph_conv0_* apply mean/rstd while loading). lengths are int64 like Batch.lengths. */
int aph_wave_stats(const float* x, const int64_t* lengths, int32_t n_utt, int32_t T,
                   double* stats_scratch, float* mean_rstd, void* stream);
int aph_wave_norm(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt,
                  int32_t T, float* out, void* stream);
/* Wav2Vec2AcousticModel.downsampled_lengths, acoustic_model.py:832-835 with
 * conv_length(use_padding=False), frontend.py:192-203: L <- floor((L - k)/s) + 1 per layer.
 * kernels/strides are device int32[n_layers]; either output may be NULL. */
int aph_frame_lengths(const int64_t* lengths, int32_t n_utt, const int32_t* kernels,
                      const int32_t* strides, int32_t n_layers, int32_t* frames32,
                      int64_t* frames64, void* stream);

/* ---- feature-extractor layer 0 and row LayerNorm --------------------------- */
/* Conv1d(1,512,k=10,s=5,bias) + LayerNorm(512) + GELU (HF:275-299, layer 0), reading the RAW
 * waveform and applying mean_rstd on the fly (NULL = no normalisation). Output bf16
 * channels-last [n_utt][L0][512], L0 = (T-10)/5+1; with skip_padded_frames != 0 frames that
 * would read padding are not computed (their rows keep stale, finite data).
 * w is the Conv1d weight [512][1][10] (fp32), gamma/beta the LayerNorm affine. */
int aph_conv0_ln_gelu(const float* x, const int64_t* lengths, const float* mean_rstd,
                      int32_t n_utt, int32_t T, const float* w, const float* bias,
                      const float* gamma, const float* beta, float eps,
                      int32_t skip_padded_frames, void* out_bf16, void* stream);
/* feat_extract_norm="group" variant (HF:302-323): conv (bias may be NULL) + GroupNorm(512,512)
 * whose statistics run over the whole padded time axis + GELU.
 * raw_scratch: fp32 [n_utt][L0][512]; stats_scratch: fp64 [n_utt][512][2]. */
int aph_conv0_gn_gelu(const float* x, const int64_t* lengths, const float* mean_rstd,
                      int32_t n_utt, int32_t T, const float* w, const float* bias,
                      const float* gamma, const float* beta, float eps, float* raw_scratch,
                      double* stats_scratch, void* out_bf16, void* stream);
/* nn.LayerNorm over the last axis (+ optional GELU), one warp per row; cols in {512, 1024}.
 * Call sites: conv layers 1-6 (HF:290-299), feature projection (HF:429-431), encoder layers
 * (HF:766-767 ×2 per layer), final encoder norm (HF:792). in: bf16 or fp32. */
int aph_layernorm_rows(const void* in, int32_t in_is_f32, int64_t ld_in, int64_t rows,
                       int32_t cols, const float* gamma, const float* beta, float eps,
                       int32_t gelu, void* out_bf16, int64_t ld_bf16, float* out_f32,
                       int64_t ld_f32, void* stream);

/* ---- classifier heads -------------------------------------------------------- */
/* EmbeddingCompositionLayer.forward, acoustic_model.py:219-232: row 0 = blank embedding
 * weight[0]; row 1+v = sum_f weight[tfi[v][f] + category_offsets[f]] (EmbeddingBag "sum").
 * category_offsets NULL = tfi already offset (the training table). Rows up to rows_out are
 * zero-filled (GEMM padding). err_flag is set to 1 on an out-of-range category. */
int aph_compose_embeddings(const float* weight, int32_t n_categories, int32_t embedding_size,
                           const int64_t* tfi, const int64_t* category_offsets,
                           int32_t n_phonemes, int32_t n_features, int32_t rows_out,
                           void* out_bf16, float* out_f32, int32_t* err_flag, void* stream);
/* log_softmax(logits, -1) for many narrow heads in one launch (acoustic_model.py:1051-1052
 * via estimator.py:1041-1045; loss_functions.py:27). logits fp32 [rows][ld]; head h reads
 * columns [col_off[h], col_off[h]+width[h]) (all inside [col_lo, col_lo+col_span)) and writes a
 * contiguous fp32 [rows][width[h]] block at out + out_off[h]. Optional per-frame argmax
 * (lowest index on ties) and max log-prob, laid out [n_heads][rows], feed greedy decoding. */
int aph_log_softmax_heads(const float* logits, int64_t ld, int64_t rows, int32_t col_lo,
                          int32_t col_span, const int32_t* col_off, const int32_t* width,
                          const int64_t* out_off, int32_t n_heads, float* out,
                          int32_t* argmax_out, float* maxlp_out, void* stream);
/* Same for one wide head (large inventories): one warp per frame, row staged in smem. */
int aph_log_softmax_wide(const float* logits, int64_t ld, int64_t rows, int32_t width,
                         float* out, int64_t ld_out, int32_t* argmax_out, float* maxlp_out,
                         void* stream);
/* HierarchicalProjection.forward dependency inputs, acoustic_model.py:497-514:
 * softmax(logits[..., skip:]) of each dependency written as bf16 into columns
 * [dst_col[d], ...) of the next classifier's input matrix. */
int aph_dependency_softmax(const float* logits, int64_t ld, int64_t rows, const int32_t* col_off,
                           const int32_t* width, const int32_t* dst_col, int32_t n_deps,
                           int32_t skip, void* dst_bf16, int64_t ld_dst, void* stream);

/* ---- greedy CTC decoding ------------------------------------------------------ */
/* torch.max(log_emissions, -1), predictions.py:195. x fp32 [rows][ld]. */
int aph_argmax_rows(const float* x, int64_t ld, int64_t rows, int32_t width, int32_t* argmax_out,
                    float* max_out, void* stream);
/* GreedyCTCDecoder.__call__, predictions.py:196-206, for n_seq = heads*n_utt sequences of T
 * frames (sequence s belongs to utterance s % n_utt): collapse repeats, drop `blank`,
 * 1-based run-start timesteps, score = sum of max log-probs over valid frames.
 * tokens/timesteps: int32 [n_seq][T] (first counts[s] entries valid). */
int aph_ctc_greedy_collapse(const int32_t* argmax_in, const float* maxlp_in,
                            const int32_t* lengths, int32_t n_utt, int32_t T, int32_t n_seq,
                            int32_t blank, int32_t* tokens, int32_t* timesteps, int32_t* counts,
                            float* scores, void* stream);

/* Dense packing of the collapse results for the device-to-host copy: offsets[n_seq+1] = exclusive scan of
 * counts; packed_tokens / packed_timesteps (capacity n_seq*T) hold hypothesis s at [offsets[s], offsets[s+1]). */
int aph_ctc_pack_hypotheses(const int32_t* tokens, const int32_t* timesteps, const int32_t* counts,
                            int32_t n_seq, int32_t T, int32_t* offsets, int32_t* packed_tokens,
                            int32_t* packed_timesteps, void* stream);

/* ---- weight packing (once per weight version) ---------------------------------- */
/* fp32 -> bf16 (nn.Linear weights are already the K-major B operand). */
int aph_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* Same for a row-strided matrix (keeps a hidden state as a bf16 classifier input block). */
int aph_cast_bf16_2d(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst,
                     int64_t rows, int32_t cols, void* stream);
/* Conv1d weight [O][C][k] (HF:281-287) -> bf16 [O][k][C] for the channels-last implicit GEMM. */
int aph_pack_conv_weight(const float* src, void* dst_bf16, int32_t out_channels,
                         int32_t in_channels, int32_t kernel, void* stream);
/* weight_norm(dim=2) of the positional conv (HF:326-350): w = g[j] * v / ||v[:,:,j]||, packed
 * to bf16 [O][k][Cg]. weight_g: parametrizations.weight.original0 [1][1][k]; weight_v:
 * original1 [O][Cg][k]; tap_scale_scratch: fp32 [k]. */
int aph_pack_posconv_weight(const float* weight_g, const float* weight_v, void* dst_bf16,
                            float* tap_scale_scratch, int32_t out_channels,
                            int32_t group_channels, int32_t kernel, void* stream);

/* LayerNorm folded into the nn.Linear that follows it (HF:735-736 layer_norm -> q/k/v_proj, HF:749-750 final_layer_norm ->
 * intermediate_dense): out_weight[n][k] = bf16(weight[n][k] * gamma[k]), out_colsum[n] = sum_k out_weight[n][k],
 * out_bias[n] = bias[n] + sum_k weight[n][k] * beta[k].  The GEMM consumes them through aph_gemm_args.ln_* (ABI 5). */
int aph_fold_layernorm_linear(const float* weight /*[n][k]*/, const float* bias /*[n] or NULL*/, const float* gamma, const float* beta,
                              int32_t n, int32_t k, void* out_weight_bf16, float* out_colsum, float* out_bias, void* stream);

/* ---- multi-head CTC loss -------------------------------------------------------- */
/* CTCWrapper (loss_functions.py:19-27): nn.CTCLoss(reduction="sum", zero_infinity=True) on
 * log_softmax(logits); blank = 0; called once per classifier head (estimator.py:721-734).
 * Here ONE launch covers every (head, utterance) pair. A head is described by: */
typedef struct aph_ctc_head {
  const float* log_probs;       /* element (n, t, k) at log_probs[n*stride_n + t*stride_t + k]   */
  float* grad;                  /* d loss / d LOGITS, same addressing; NULL = no gradient        */
  int64_t stride_t;
  int64_t stride_n;
  int32_t n_classes;
  int32_t s_pad;                /* aph_ctc_states_pad(max label length over all heads)           */
  const int64_t* labels;        /* int64 [n_utt][label_stride], zero padded, values >= 1          */
  int64_t label_stride;
  const int64_t* label_lengths; /* int64 [n_utt]                                                  */
  int64_t alpha_offset;         /* float offset of this head's [n_utt][T][s_pad] block in alpha_ws */
} aph_ctc_head;

/* Padded number of CTC states per frame used for the alpha workspace (negative = unsupported). */
int aph_ctc_states_pad(int32_t max_label_len);
/* Forward: nll_out[h][n] = -log p(labels | log_probs) (inf when infeasible); loss_out[h] =
 * sum_n nll with infinities zeroed (may be NULL). alpha_ws (may be NULL when no gradient is
 * needed) receives the forward variables for aph_ctc_backward. heads_host is a HOST array (the descriptors
 * travel in the kernel parameters; the pointers inside are device pointers). */
int aph_ctc_forward(const aph_ctc_head* heads_host, int32_t n_heads, int32_t n_utt, int32_t T,
                    int32_t max_label_len, const int64_t* input_lengths, float* alpha_ws,
                    float* nll_out, float* loss_out, void* stream);
/* Backward: writes grad (w.r.t. the logits that produced log_probs) for every head with a
 * non-NULL grad pointer: grad_scale[h] * (softmax - state occupancies), zero for padded
 * frames and for pairs whose loss was infinite. alpha_ws is CONSUMED: the block-per-pair
 * recursions overwrite the forward variables with the state occupancies (one backward call per
 * forward call, with the same n_heads / n_utt / max_label_len: they select the kernels). */
int aph_ctc_backward(const aph_ctc_head* heads_host, int32_t n_heads, int32_t n_utt, int32_t T, int32_t max_label_len,
                     const int64_t* input_lengths, const float* alpha_ws, const float* nll,
                     const float* grad_scale, void* stream);

/* ---- training-side kernels of the classifier heads ------------------------------------ */
/* AllophoneMapping.map_allophones + _multiply_allophone_matrix (acoustic_model.py:75-87, 142-159):
 * out[n][t][q] = max_p (mask[lang][p][q] ? finfo(float32).min : logits[n][t][p] * matrices[lang][p][q]).
 * The allowed (p, q) pairs of every language are given as a CSR list over q
 * (csr_offsets[lang*n_phonemes + q] .. [+1] indexes csr_phones). logits element (n,t,p) is at
 * logits[n*stride_n + t*stride_t + p]; out/argmax are contiguous [n_utt][T][n_phonemes]. */
int aph_allophone_forward(const float* logits, int64_t stride_n, int64_t stride_t, int32_t n_utt,
                          int32_t T, int32_t n_phones, int32_t n_phonemes, const float* matrices,
                          const int32_t* csr_offsets, const int32_t* csr_phones,
                          const int64_t* language_ids, float* out, int32_t* argmax_out, void* stream);
/* Its backward: grad_logits (contiguous [n_utt][T][n_phones], pre-zeroed) and grad_matrices
 * (same shape as matrices, accumulated) receive the gradient of the winning phone. */
int aph_allophone_backward(const float* grad_out, const int32_t* argmax_in, const float* logits,
                           int64_t stride_n, int64_t stride_t, int32_t n_utt, int32_t T,
                           int32_t n_phones, int32_t n_phonemes, const float* matrices,
                           const int64_t* language_ids, float* grad_logits, float* grad_matrices,
                           void* stream);
/* [rows][cols] (fp32 or bf16) -> bf16 [cols][rows_padded] (zero padded): turns the frame axis into
 * the contiguous K axis of the weight-gradient GEMMs dW = dY^T X (autograd of nn.Linear). */
int aph_transpose_cast_bf16(const void* in, int32_t in_is_f32, int64_t ld_in, int64_t rows,
                            int32_t cols, void* out_bf16, int64_t ld_out, int64_t rows_padded,
                            void* stream);
/* out[c] = sum_r in[r][c] (bias gradients). */
int aph_colsum_f32(const float* in, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream);
int aph_colsum_bf16(const void* in_bf16, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream);
/* Backward of nn.LayerNorm over the last axis (cols in {512, 1024}); statistics are recomputed from x.
 *   dx[r] = (dx_resid ? dx_resid[r] : 0) + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
 *   dgamma[c] = sum_r dy * xhat, dbeta[c] = sum_r dy (either may be NULL).  x, dy: fp32 or bf16. dx may alias dx_resid. */
int aph_layernorm_backward(const void* x, int32_t x_is_f32, int64_t ld_x, const void* dy,
                           int32_t dy_is_f32, int64_t ld_dy, int64_t rows, int32_t cols,
                           const float* gamma, float eps, const float* dx_resid, int64_t ld_resid,
                           float* dx, int64_t ld_dx, float* dgamma, float* dbeta, void* stream);
/* Backward of `hidden_states[~mask] = 0` (HF:753-756): zero the rows t >= lengths[row / len_period]. */
int aph_mask_rows_f32(float* x, int64_t ld, int64_t rows, int32_t cols, const int32_t* lengths,
                      int32_t len_period, void* stream);
/* dst[r][c] += src[r][c] (gradient of a hidden state that is also a classifier input, OUTPUT_i). */
int aph_add_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows,
                   int32_t cols, void* stream);
/* out = bf16(dy * gelu'(pre)): backward of the GELU after the positional conv (HF:353-368). */
int aph_gelu_backward_bf16(const float* dy, int64_t ld_dy, const void* pre_bf16, int64_t ld_pre,
                           int64_t rows, int32_t cols, void* out_bf16, int64_t ld_out, void* stream);
/* Positional conv backward (weight_norm(dim=2) grouped Conv1d, HF:326-350).
 * aph_pack_posconv_weight_dgrad: B operand of the data-gradient sliding-tap GEMM (tap_pad = k/2 - 1),
 *   dst[g*Cg + ci][j*Cg + co] = w[g*Cg + co][ci][k-1-j]; tap_scratch: fp32 [2*k].
 * aph_posconv_weight_backward: raw = output of the APH_GEMM_DIAG_TAPS GEMM, fp32 [k][O][256] -> gradients of
 *   weight_g [1][1][k] and weight_v [O][Cg][k]; tap_scratch: fp32 [3*k]. */
int aph_pack_posconv_weight_dgrad(const float* weight_g, const float* weight_v, void* dst_bf16,
                                  float* tap_scratch, int32_t out_channels, int32_t group_channels,
                                  int32_t kernel, void* stream);
int aph_posconv_weight_backward(const float* raw, const float* weight_g, const float* weight_v,
                                float* tap_scratch, int32_t out_channels, int32_t group_channels,
                                int32_t kernel, float* grad_g, float* grad_v, void* stream);
/* The same with raw = fp32 [k][O][block_width]: row o holds the products of output channel o with the block_width input channels of
 * its diagonal block (block_width a multiple of the group width that divides O).  256 = what APH_GEMM_DIAG_TAPS writes; O = one full
 * [O][O] weight-gradient GEMM per tap (b_k_shift = tap - k/2), used where the groups do not tile 256 channels (wav2vec2-base:
 * 16 groups of 48). */
int aph_posconv_weight_backward_blocks(const float* raw, const float* weight_g, const float* weight_v,
                                       float* tap_scratch, int32_t out_channels, int32_t group_channels,
                                       int32_t kernel, int32_t block_width, float* grad_g, float* grad_v, void* stream);
/* Backward of EmbeddingCompositionLayer's EmbeddingBag("sum") (acoustic_model.py:208, 225-232):
 * grad_rows[0] -> category 0 (blank), grad_rows[1+v] -> every category tfi[v][f] + offsets[f]. */
int aph_embedding_bag_backward(const float* grad_rows, int64_t ld, int32_t n_phonemes,
                               int32_t n_features, int32_t embedding_size, const int64_t* tfi,
                               const int64_t* category_offsets, float* grad_weight, void* stream);
/* One classifier head's [rows][width] fp32 block inside a row-major matrix (ptr = its first element). */
typedef struct aph_head_block {
  void* ptr;
  int64_t ld; /* row stride in elements */
  int32_t width;
  int32_t rows; /* 0 = the row count of the call; > 0 = this block's own (the weight blocks of the level matrices) */
} aph_head_block;
/* dst_b = src_b (accumulate == 0) or dst_b += src_b for every block b in ONE launch: the per-head logits handed to autograd
 * (acoustic_model.py:395-416 returns one tensor per classifier) and the logits gradients collected into the level's gradient
 * matrix on the way back; also the per-step assembly of a level's weight / bias matrix from the classifiers' nn.Linear parameters
 * (blocks with their own row counts). src_host / dst_host are HOST arrays (descriptors travel in the kernel parameters). */
int aph_copy_head_blocks(const aph_head_block* src_host, const aph_head_block* dst_host, int32_t n_blocks, int64_t rows,
                         int32_t accumulate, void* stream);
/* functional.log_softmax(x, -1) of every block in ONE launch (loss_functions.py:26: CTCWrapper applies it per head). */
int aph_log_softmax_head_blocks(const aph_head_block* src_host, const aph_head_block* dst_host, int32_t n_blocks, int64_t rows,
                                void* stream);
/* Backward of the dependency softmax (acoustic_model.py:497-514): for dependency d,
 * grad_logits[:, dst_col[d]+skip : ...] += p * (dp - sum(p*dp)) with p the bf16 probabilities stored
 * in x[:, x_col[d] : ...] and dp = grad_x[:, x_col[d] : ...]. */
int aph_softmax_backward_cols(const float* grad_x, int64_t ld_gx, const void* x_bf16, int64_t ld_x,
                              int64_t rows, const int32_t* x_col, const int32_t* width,
                              const int32_t* dst_col, int32_t n_deps, int32_t skip,
                              float* grad_logits, int64_t ld_gl, void* stream);

/* ---- optimiser step (estimator.py:778-791; config.py:316-335) ------------------------------ */
/* Multi-tensor kernels: `tensors_host` / `sizes_host` are HOST arrays of device pointers and element counts (the
 * descriptors travel in the kernel parameters, up to 48 tensors per launch); all tensors are contiguous fp32.
 * aph_multi_tensor_sumsq: *out (device double) = sum over all tensors of x^2 (nn.utils.clip_grad_norm_, norm 2).
 * aph_multi_tensor_scale: x *= min(1, max_norm / (sqrt(*sumsq) + 1e-6)) in place (the clip itself).
 * aph_multi_tensor_adam: torch.optim.Adam(betas, eps, weight_decay as L2 added to the gradient) for step `step`
 *   (1-based); tensors_host is [n][4] = {param, grad, exp_avg, exp_avg_sq}; shadow_bf16_host[n] (entries may be
 *   NULL) receive the updated parameter rounded to bf16 (the GEMM operand copy); with clip_sumsq != NULL the
 *   gradients are scaled by the clip coefficient on the fly instead of in place. */
int aph_multi_tensor_sumsq(void* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors,
                           double* out, void* stream);
int aph_multi_tensor_scale(void* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors,
                           const double* sumsq, float max_norm, void* stream);
int aph_multi_tensor_adam(void* const* tensors_host, void* const* shadow_bf16_host,
                          const int64_t* sizes_host, int32_t n_tensors, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int64_t step,
                          const double* clip_sumsq, float max_norm, void* stream);
/* torch.optim.SGD(lr, momentum, weight_decay; dampening 0, no Nesterov) (config.py:300-312), same multi-tensor form:
 * tensors_host [n][3] = param, grad, momentum buffer (NULL without momentum); first_step: buffers are initialised. */
int aph_multi_tensor_sgd(void* const* tensors_host, void* const* shadow_bf16_host, const int64_t* sizes_host,
                         int32_t n_tensors, float lr, float momentum, float weight_decay, int32_t first_step,
                         const double* clip_sumsq, float max_norm, void* stream);

/* ---- edit distance (host) --------------------------------------------------------- */
/* Batched replacement of the Rust extension `allophant.phonemes` (src/edit_distance.rs):
 * levensthein (70-96) and levensthein_statistics (601-608 -> 372-481, uniform costs 483-496)
 * over int64 symbol ids. Pair p compares expected[expected_offsets[p]:expected_offsets[p+1]]
 * with actual[actual_offsets[p]:actual_offsets[p+1]]; statistics rows are
 * {insertions, deletions, substitutions, correct}. ALL POINTERS ARE HOST POINTERS; the work is
 * spread over n_threads host threads (<= 0: all cores). Either output may be NULL. */
int aph_edit_statistics_batch(const int64_t* expected_host, const int64_t* expected_offsets_host,
                              const int64_t* actual_host, const int64_t* actual_offsets_host,
                              int64_t n_pairs, uint64_t* statistics_host,
                              uint64_t* distances_host, int32_t n_threads);
/* EditStatistics.word_error_rate (src/edit_distance.rs:311-317): (S+D+I)/(S+D+C) in f32. */
float aph_word_error_rate(uint64_t insertions, uint64_t deletions, uint64_t substitutions,
                          uint64_t correct);

/* ---- training of the wav2vec2 convolutional feature extractor (freeze_feature_encoder = false / UnfreezeSchedule; ----
 * ---- HF modeling_wav2vec2.py:275-323, 382-419, layer-norm variant) -------------------------------------------------- */
/* conv(1 -> 512, k 10, s 5) of the normalised waveform BEFORE LayerNorm, bf16 [n_utt][L0][512] (kept for the backward). */
int aph_conv0_raw_bf16(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt, int32_t T,
                       const float* w, const float* bias, void* out_bf16, void* stream);
/* Backward of LayerNorm(512) -> GELU from the kept pre-LayerNorm conv output x (bf16 [rows][512]) and the gradient of the
 * GELU output d_out (fp32 [rows][ld_d]): dx bf16 [rows][512]; dgamma, dbeta and (optional) dbias = column sums of dx are
 * OVERWRITTEN. */
int aph_ln_gelu_backward_512(const void* x_bf16, const float* d_out, int64_t ld_d, int64_t rows, const float* gamma,
                             const float* beta, float eps, void* dx_bf16, float* dgamma, float* dbeta, float* dbias,
                             void* stream);
/* dW0 fp32 [512][10] = sum_(n,t) dY[n][t][o] * normalised_waveform[n][5 t + j]. */
int aph_conv0_weight_backward(const void* dy_bf16, const float* x, const int64_t* lengths, const float* mean_rstd,
                              int32_t n_utt, int32_t T, float* dw, void* stream);

/* ---- from-scratch pre-LN transformer acoustic model (acoustic_model.py:34-69, 564-759; frontend.py; padding.py) */
/* nn.LayerNorm over the last axis of fp32 x [rows][ld_x], any width; gamma/beta NULL = elementwise_affine=False. */
int aph_layernorm_any(const float* x, int64_t ld_x, int64_t rows, int32_t cols, const float* gamma,
                      const float* beta, float eps, float* out_f32, int64_t ld_f32, void* out_bf16,
                      int64_t ld_bf16, void* stream);
/* Backward of aph_layernorm_any: dx = LN'(dy) (+ resid); dgamma / dbeta are ACCUMULATED (pre-zero them). */
int aph_layernorm_any_backward(const float* x, int64_t ld_x, const float* dy, int64_t ld_dy, int64_t rows,
                               int32_t cols, const float* gamma, float eps, const float* resid, int64_t ld_resid,
                               float* dx, int64_t ld_dx, float* dgamma, float* dbeta, void* stream);
/* Backward of aph_glu_rows: dy bf16 [rows][2*out_channels] = (d_out * sigmoid(g) | d_out * a * sigmoid'(g)). */
int aph_glu_backward_bf16(const float* y, int64_t ld_y, const float* d_out, int64_t ld_d, int64_t rows,
                          int32_t out_channels, void* dy_bf16, int64_t ld_dy, void* stream);
/* Backward of (mask -> aph_reflect_pad_bf16 -> strided Conv1d) w.r.t. the stage input from the per-window gradients
 * d_cols fp32 [n_utt*out_len][kernel*channels] (= dY W): col2im overlap-add, the reflections folded back (the left one into
 * utterance 0, like the forward reads it), masked frames zero.  d_x fp32 [n_utt][length][ld_dx]. */
int aph_conv_input_backward(const float* d_cols, const int32_t* lengths, int32_t n_utt, int32_t length,
                            int32_t channels, int32_t out_len, int32_t kernel, int32_t stride, int32_t left,
                            int32_t right, int32_t reflect, float* d_x, int64_t ld_dx, void* stream);
/* d <- d * act'(.) decided from the activation OUTPUT y; kind 2 = ReLU, 3 = LeakyReLU(0.01); optional bf16 copy. */
int aph_activation_backward(float* d, int64_t ld_d, const float* y, int64_t ld_y, int64_t rows, int32_t cols,
                            int32_t kind, void* out_bf16, int64_t ld_bf16, void* stream);
/* SinusoidalPositionEmbeddings.forward (acoustic_model.py:58-69): x[n][t][c] += sin|cos(t * bases[c]). */
int aph_add_sinusoidal(float* x, int64_t ld, int32_t n_utt, int32_t seq, int32_t cols, const float* bases,
                       void* stream);

/* Time layer of a classifier head: nn.MultiheadAttention over the frames of an utterance on the head's projected classes
 * (ProjectingMultiheadAttention, acoustic_model.py:237-268, used at 406-413) with the key-padding mask of the frame counts.
 * qkv fp32 [n_utt*T][ld] = in_proj output (q | k | v, `heads * head_dim` columns each); ctx bf16 [n_utt*T][ld_ctx] receives the
 * heads' outputs, the operand of out_proj.  Any head_dim (CUDA cores; the encoder's tcgen05 attention needs 64). */
int aph_attention_small(const float* qkv, int64_t ld, void* ctx_bf16, int64_t ld_ctx, const int32_t* lengths, int32_t n_utt,
                        int32_t heads, int32_t T, int32_t head_dim, void* stream);
/* features [N][F][L] fp32 -> channels-last [N][L][ld_out] fp32 (the permute of acoustic_model.py:672). */
int aph_transpose_nfl(const float* in, int32_t n_utt, int32_t features, int32_t length, float* out,
                      int64_t ld_out, void* stream);
/* LengthWrapper masking + VariableLengthReflectPad (frontend.py:63-74, padding.py:41-53) on channels-last fp32
 * x [N][length][ld_x] -> bf16 [N][length+left+right][channels]; reflect = 0: zero padding. */
int aph_reflect_pad_bf16(const float* x, int64_t ld_x, const int32_t* lengths, int32_t n_utt, int32_t length,
                         int32_t channels, int32_t left, int32_t right, int32_t reflect, void* out_bf16,
                         void* stream);
/* functional.glu over channels (frontend.py:136): out[r][c] = y[r][c] * sigmoid(y[r][out_channels + c]). */
int aph_glu_rows(const float* y, int64_t ld_y, int64_t rows, int32_t out_channels, float* out, int64_t ld_out,
                 void* stream);

/* ---- rest of the Rust extension allophant.phonemes (host; ALL POINTERS ARE HOST POINTERS) ------------------ */
/* levensthein_matrix (src/edit_distance.rs:261-269): (m+1) x (n+1) fp32 uniform-cost matrix of int64 symbol ids. */
int aph_edit_matrix(const int64_t* a, int64_t m, const int64_t* b, int64_t n, float* matrix_out);
/* levensthein_operations (src/edit_distance.rs:116-218, 271-280): first best path, (action, i, j) triples in forward
 * order (action 0 insertion / 1 deletion / 2 substitution), at most m + n; returns their number or a negative error. */
int64_t aph_edit_operations(const int64_t* a, int64_t m, const int64_t* b, int64_t n, int64_t* ops_out,
                            float* final_cost);
/* PropertyWeighting (src/edit_distance.rs:497-598): weighted edit distance with sub_cost[i*n + j] = substitution cost of
 * (a_i, b_j); mode 0 -> matrix_out (m+1)x(n+1), 1 -> ops_out (returns their number), 2 -> stats_out[4]. */
int64_t aph_edit_weighted(int64_t m, int64_t n, const float* sub_cost, float insertion_cost, float deletion_cost,
                          int32_t mode, float* matrix_out, int64_t* ops_out, uint64_t* stats_out,
                          float* final_cost);
/* IpaSegmenter (src/ipa_segmenter.rs): leftmost-longest non-overlapping vocabulary matches over UTF-8 bytes. */
void* aph_segmenter_create(const char* blob, const int64_t* offsets, int64_t n);
void aph_segmenter_free(void* handle);
int64_t aph_segmenter_find(const void* handle, const char* text, int64_t text_len, int64_t* bounds_out,
                           int64_t max_matches);

/* CTC beam search (predictions.py:210-226 -> torchaudio ctc_decoder(lexicon=None, lm=None, log_add=True) -> flashlight-text
 * LexiconFreeDecoder + ZeroLM, restated): log_emissions fp32 [n_seq][t_max][classes], lengths int64 [n_seq]; per sequence
 * the nbest best token sequences: tokens / timesteps int64 [n_seq][nbest][t_max], counts int64 [n_seq][nbest] (-1 = none),
 * scores fp64 [n_seq][nbest].  beam_size_token <= 0: all classes.  HOST pointers; n_threads <= 0: all cores. */
int aph_ctc_beam_decode(const float* log_emissions, const int64_t* lengths, int64_t n_seq, int64_t t_max,
                        int32_t classes, int32_t blank, int32_t beam_size, int32_t beam_size_token,
                        double beam_threshold, int32_t nbest, int32_t log_add, int64_t* tokens_out,
                        int64_t* timesteps_out, int64_t* counts_out, double* scores_out, int32_t n_threads);

/* ---- host feeding (batching.py:171-215) -------------------------------------------------------- */
/* rnn.pad_sequence of the utterances of a batch: n fp32 arrays of lengths_host[i] samples -> zero-padded
 * [n][max_len] (normally a pinned staging buffer), spread over n_threads host threads (<= 0: all cores).
 * ALL POINTERS ARE HOST POINTERS. */
int aph_collate_pad_f32(const float* const* utterances_host, const int64_t* lengths_host, int64_t n,
                        int64_t max_len, float* dst_host, int32_t n_threads);

#ifdef __cplusplus
}
#endif

#endif /* ALLOPHANT_B200_H_ */
