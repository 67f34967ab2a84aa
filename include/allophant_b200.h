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

#define APH_ABI_VERSION 1

/* ---- library ------------------------------------------------------------ */
int aph_abi_version(void);
const char* aph_last_error(void);
/* Number of kernels launched by this library since the last reset (bench.py
 * reports it as `gpu_launches`). */
int64_t aph_launch_count(void);
void aph_reset_launch_count(void);

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
#define APH_EPI_STORE 0    /* generic epilogue described above                                   */
#define APH_EPI_QKV 1      /* scatter bf16 into Q[b,h,t,64] (pre-scaled), K[b,h,t,64], Vt[b,h,64,t_v] */

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
  int32_t gelu;
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
  void* vt;
  int32_t heads;
  int32_t t_v;            /* padded key length of Vt (multiple of 8) */
  float q_scale;
} aph_gemm_args;

int aph_gemm_bf16(const aph_gemm_args* args, void* stream);

/* ---- variable-length self-attention (tcgen05, flash-style online softmax) --- */
/* Replaces the SDPA call of Wav2Vec2Attention (HF:466-549) and the dense
 * additive mask of HF:758-762: keys t >= lengths[b] get probability 0.
 *   q, k : bf16 [n_utt*heads][T][64]   (q already scaled by head_dim^-0.5)
 *   vt   : bf16 [n_utt*heads][64][t_v] (V transposed, keys contiguous, t_v % 8 == 0,
 *          columns [T, t_v) must be finite)
 *   ctx  : bf16 [n_utt*T][heads*64]    rows of padded query tiles are left untouched
 */
int aph_attention_bf16(const void* q, const void* k, const void* vt, void* ctx,
                       const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                       int32_t t_v, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* ALLOPHANT_B200_H_ */
