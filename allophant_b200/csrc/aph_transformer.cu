// Small kernels of the from-scratch pre-LN transformer acoustic model (acoustic_model.py:34-69, 564-759; frontend.py:98-276;
// padding.py:24-53): LayerNorm over an arbitrary feature width with optional affine parameters, sinusoidal position
// embeddings, [N, F, L] -> channels-last transpose, variable-length reflect padding and the gated linear unit.  All
// HBM-bound elementwise / row kernels; the GEMMs and the attention of this model are the kernels of aph_gemm.cu /
// aph_attention.cu.
#include "aph_common.cuh"

namespace aph {

// nn.LayerNorm over the last axis, any width, elementwise_affine optional; warp per row, fp32 two-pass statistics.
__global__ void __launch_bounds__(256) layernorm_any_kernel(const float* __restrict__ x, long long ld_x, long long rows, int cols,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                            float* out_f32, long long ld_f32, __nv_bfloat16* out_bf16, long long ld_bf16) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = x + row * ld_x;
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += src[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(cols);
  float sq = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float d = src[c] - mean;
    sq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(cols) + eps);
  for (int c = lane; c < cols; c += 32) {
    float v = (src[c] - mean) * rstd;
    if (gamma != nullptr) v = v * gamma[c] + beta[c];
    if (out_f32 != nullptr) out_f32[row * ld_f32 + c] = v;
    if (out_bf16 != nullptr) out_bf16[row * ld_bf16 + c] = __float2bfloat16(v);
  }
}

// x[n][t][c] += sin / cos(t * bases[c]) (even / odd c): SinusoidalPositionEmbeddings.forward, acoustic_model.py:58-69
__global__ void __launch_bounds__(256) add_sinusoidal_kernel(float* x, long long ld, int seq, int cols, long long rows,
                                                             const float* __restrict__ bases) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cols;
    const int c = static_cast<int>(i - row * cols);
    const float angle = static_cast<float>(row % seq) * bases[c];
    x[row * ld + c] += (c & 1) ? cosf(angle) : sinf(angle);
  }
}

// [N][F][L] fp32 -> [N][L][F] fp32 (ld_out >= F), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_nfl_kernel(const float* __restrict__ in, int features, int length, float* __restrict__ out,
                                                            long long ld_out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int f0 = blockIdx.y * 32, l0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* src = in + static_cast<long long>(n) * features * length;
  for (int j = ty; j < 32; j += 8)
    if (f0 + j < features && l0 + tx < length) tile[j][tx] = src[static_cast<long long>(f0 + j) * length + l0 + tx];
  __syncthreads();
  float* dst = out + static_cast<long long>(n) * length * ld_out;
  for (int j = ty; j < 32; j += 8)
    if (l0 + j < length && f0 + tx < features) dst[static_cast<long long>(l0 + j) * ld_out + f0 + tx] = tile[tx][j];
}

// LengthWrapper masking + VariableLengthReflectPad (frontend.py:63-74, padding.py:41-53) on channels-last activations:
// out[n][p][c], p in [0, L + left + right); q = p - left:
//   q < 0: xm0[left - p]     0 <= q < len: x[q]      len <= q < len + right: x[len - 2 - (q - len)]      else 0
// xm0 is the (length-masked) FIRST utterance of the batch for every n: the reference gathers the left reflection with an
// index tensor of batch size 1 (padding.py:44-46: `_left_pad_indices.repeat(1, feature_size, 1)`), which reads batch
// element 0 only and broadcasts it over the batch; reproduced here because parity is defined by the reference's
// results.  Output bf16 (the conv GEMM's A operand).
__global__ void __launch_bounds__(256) reflect_pad_kernel(const float* __restrict__ x, long long ld_x, const int* __restrict__ lengths,
                                                          int length, int channels, int left, int right, int reflect,
                                                          __nv_bfloat16* __restrict__ out) {
  const int padded = length + left + right;
  const long long total = static_cast<long long>(gridDim.y) * padded * channels;
  (void)total;
  const int n = blockIdx.y;
  const int len = min(lengths[n], length);
  const float* src = x + static_cast<long long>(n) * length * ld_x;
  __nv_bfloat16* dst = out + static_cast<long long>(n) * padded * channels;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < static_cast<long long>(padded) * channels; i += 256ll * gridDim.x) {
    const int p = static_cast<int>(i / channels);
    const int c = static_cast<int>(i - static_cast<long long>(p) * channels);
    const int q = p - left;
    int t = -1;
    if (q < 0) {
      if (reflect) {  // left reflection of utterance 0, for every utterance (see above)
        const int t0 = left - p;
        const float v0 = t0 < min(lengths[0], length) ? x[static_cast<long long>(t0) * ld_x + c] : 0.f;
        dst[i] = __float2bfloat16(v0);
        continue;
      }
    } else if (q < len) {
      t = q;
    } else if (reflect && q < len + right) {
      t = len - 2 - (q - len);
    }
    const float v = (t >= 0 && t < len) ? src[static_cast<long long>(t) * ld_x + c] : 0.f;
    dst[i] = __float2bfloat16(v);
  }
}

// functional.glu over the channel axis of a channels-last matrix: out[r][c] = y[r][c] * sigmoid(y[r][O + c])
__global__ void __launch_bounds__(256) glu_rows_kernel(const float* __restrict__ y, long long ld_y, long long rows, int out_channels,
                                                       float* __restrict__ out, long long ld_out) {
  const long long total = rows * out_channels;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / out_channels;
    const int c = static_cast<int>(i - row * out_channels);
    const float a = y[row * ld_y + c], g = y[row * ld_y + out_channels + c];
    out[row * ld_out + c] = a / (1.0f + __expf(-g));
  }
}

// Backward of nn.LayerNorm over the last axis, any width, optional affine parameters; warp per row, statistics recomputed.
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;   dx_out = dx (+ resid);  dgamma += dy * xhat, dbeta += dy
// dgamma / dbeta are ACCUMULATED with atomics (the final LayerNorm of the transformer model is applied to several layer
// outputs); they are only present for elementwise_affine models.
__global__ void __launch_bounds__(256) layernorm_any_backward_kernel(const float* __restrict__ x, long long ld_x, const float* __restrict__ dy,
                                                                     long long ld_dy, long long rows, int cols,
                                                                     const float* __restrict__ gamma, float eps, const float* resid,
                                                                     long long ld_resid, float* dx, long long ld_dx, float* dgamma,
                                                                     float* dbeta) {
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ld_x;
  const float* dyr = dy + row * ld_dy;
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(cols);
  float sq = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float d = xr[c] - mean;
    sq += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(cols) + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float g = dyr[c] * (gamma != nullptr ? gamma[c] : 1.f);
    s1 += g;
    s2 += g * (xr[c] - mean) * rstd;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float c1 = s1 / static_cast<float>(cols), c2 = s2 / static_cast<float>(cols);
  for (int c = lane; c < cols; c += 32) {
    const float xhat = (xr[c] - mean) * rstd;
    const float d = dyr[c];
    if (dx != nullptr) {
      float v = rstd * (d * (gamma != nullptr ? gamma[c] : 1.f) - c1 - xhat * c2);
      if (resid != nullptr) v += resid[row * ld_resid + c];
      dx[row * ld_dx + c] = v;
    }
    if (dgamma != nullptr) {
      atomicAdd(dgamma + c, d * xhat);
      atomicAdd(dbeta + c, d);
    }
  }
}

// Backward of functional.glu over channels: y = [a | g], out = a * sigmoid(g):
//   d_a = d_out * sigmoid(g),   d_g = d_out * a * sigmoid(g) * (1 - sigmoid(g));   written as bf16 (the operand of the conv's
//   data-gradient and weight-gradient GEMMs)
__global__ void __launch_bounds__(256) glu_backward_kernel(const float* __restrict__ y, long long ld_y, const float* __restrict__ d_out,
                                                           long long ld_d, long long rows, int out_channels,
                                                           __nv_bfloat16* __restrict__ dy, long long ld_dy) {
  const long long total = rows * out_channels;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / out_channels;
    const int c = static_cast<int>(i - row * out_channels);
    const float a = y[row * ld_y + c], g = y[row * ld_y + out_channels + c];
    const float sig = 1.0f / (1.0f + __expf(-g));
    const float d = d_out[row * ld_d + c];
    dy[row * ld_dy + c] = __float2bfloat16(d * sig);
    dy[row * ld_dy + out_channels + c] = __float2bfloat16(d * a * sig * (1.0f - sig));
  }
}

// Backward of (LengthWrapper mask -> VariableLengthReflectPad -> Conv1d) w.r.t. the stage input, from the per-window
// gradients d_cols[n][t][j*C + c] = (dY W)[n][t][j][c] of the convolution (one data-gradient GEMM):
//   d_pad[n][p][c] = sum_{j < kernel, (p - j) % stride == 0, t = (p - j) / stride < out_len} d_cols[n][t][j*C + c]      (col2im)
//   d_x[n][q][c]   = d_pad[n][q + left]                                             (the copied frame)
//                  + d_pad[n][len + left + (len - 2 - q)]   if 0 <= len - 2 - q < right   (right reflection reads x[len-2-j])
//                  + sum_n' d_pad[n'][left - q]             if n == 0 and 1 <= q <= left   (every utterance's left reflection
//                                                                                            reads utterance 0, see the forward)
// for q < len, and 0 for the masked frames q >= len.
__device__ __forceinline__ float conv_pad_gradient(const float* __restrict__ d_cols, long long n, int p, int c, int out_len, int kernel,
                                                   int stride, int channels) {
  float sum = 0.f;
  const long long width = static_cast<long long>(kernel) * channels;
  for (int j = 0; j < kernel; ++j) {
    const int r = p - j;
    if (r < 0 || r % stride != 0) continue;
    const int t = r / stride;
    if (t >= out_len) continue;
    sum += d_cols[(n * out_len + t) * width + static_cast<long long>(j) * channels + c];
  }
  return sum;
}

__global__ void __launch_bounds__(256) conv_input_backward_kernel(const float* __restrict__ d_cols, const int* __restrict__ lengths, int n_utt,
                                                                  int length, int channels, int out_len, int kernel, int stride, int left,
                                                                  int right, int reflect, float* __restrict__ d_x, long long ld_dx) {
  const int n = blockIdx.y;
  const int len = min(lengths[n], length);
  const int len0 = min(lengths[0], length);
  const long long total = static_cast<long long>(length) * channels;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int q = static_cast<int>(i / channels);
    const int c = static_cast<int>(i - static_cast<long long>(q) * channels);
    float v = 0.f;
    if (q < len) {
      v = conv_pad_gradient(d_cols, n, q + left, c, out_len, kernel, stride, channels);
      if (reflect) {
        const int j = len - 2 - q;
        if (j >= 0 && j < right) v += conv_pad_gradient(d_cols, n, len + left + j, c, out_len, kernel, stride, channels);
        if (n == 0 && q >= 1 && q <= left && q < len0) {
          for (int other = 0; other < n_utt; ++other) v += conv_pad_gradient(d_cols, other, left - q, c, out_len, kernel, stride, channels);
        }
      }
    }
    d_x[(static_cast<long long>(n) * length + q) * ld_dx + c] = v;
  }
}

// d_pre = d_out * act'(.) from the activation's OUTPUT y (kind 2 ReLU: y > 0; kind 3 LeakyReLU(0.01): y > 0 ? 1 : 0.01),
// fp32 in place + optional bf16 copy (the operand of the weight-gradient GEMM)
__global__ void __launch_bounds__(256) activation_backward_kernel(float* d, long long ld_d, const float* __restrict__ y, long long ld_y,
                                                                  long long rows, int cols, int kind, __nv_bfloat16* out_bf16,
                                                                  long long ld_bf16) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cols;
    const int c = static_cast<int>(i - row * cols);
    const float slope = y[row * ld_y + c] > 0.f ? 1.f : (kind == 3 ? 0.01f : 0.f);
    const float v = d[row * ld_d + c] * slope;
    d[row * ld_d + c] = v;
    if (out_bf16 != nullptr) out_bf16[row * ld_bf16 + c] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------
// Self-attention over TIME for narrow sequences of vectors: the time layer of a classifier head
// (ProjectingMultiheadAttention, acoustic_model.py:237-268: nn.MultiheadAttention over the head's projected classes, a few
// channels per head, with the key-padding mask of the utterance lengths).  head_dim is arbitrary (the tcgen05 kernel of the
// encoder needs 64), so this runs on the CUDA cores: one warp per query frame, scores of all keys in shared memory (two passes:
// softmax statistics, then the weighted sum with the lanes spread over the head's channels).
//   qkv  fp32 [n_utt*T][ld]: q at column 0, k at column hidden, v at column 2*hidden (nn.MultiheadAttention's in_proj order)
//   ctx  bf16 [n_utt*T][ld_ctx], columns [0, hidden); rows of padded frames are written too (computed over the valid keys)
// ---------------------------------------------------------------------------
constexpr int kSmallAttWarps = 8;

__global__ void __launch_bounds__(kSmallAttWarps * 32) attention_small_kernel(const float* __restrict__ qkv, long long ld,
                                                                             __nv_bfloat16* __restrict__ ctx, long long ld_ctx,
                                                                             const int* __restrict__ lengths, int T, int heads,
                                                                             int head_dim, float scale) {
  extern __shared__ float small_att_smem[];  // [warps][T] scores, then [warps][head_dim] the query
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = blockIdx.y;
  const int n = nh / heads, h = nh - n * heads;
  const int q_index = blockIdx.x * kSmallAttWarps + warp;
  if (q_index >= T) return;
  const int hidden = heads * head_dim;
  int len = lengths[n];
  len = len < T ? (len < 0 ? 0 : len) : T;
  float* scores = small_att_smem + static_cast<long long>(warp) * T;
  float* query = small_att_smem + static_cast<long long>(kSmallAttWarps) * T + warp * head_dim;
  const float* base = qkv + static_cast<long long>(n) * T * ld + h * head_dim;
  for (int j = lane; j < head_dim; j += 32) query[j] = base[static_cast<long long>(q_index) * ld + j] * scale;
  __syncwarp();
  float m = -INFINITY;
  for (int key = lane; key < len; key += 32) {
    const float* k_row = base + static_cast<long long>(key) * ld + hidden;
    float dot = 0.f;
    for (int j = 0; j < head_dim; ++j) dot = fmaf(query[j], k_row[j], dot);
    scores[key] = dot;
    m = fmaxf(m, dot);
  }
  m = warp_max(m);
  float total = 0.f;
  for (int key = lane; key < len; key += 32) {
    const float e = __expf(scores[key] - m);
    scores[key] = e;
    total += e;
  }
  total = warp_sum(total);
  __syncwarp();
  const float inv = len > 0 ? 1.0f / total : 0.f;
  __nv_bfloat16* out = ctx + (static_cast<long long>(n) * T + q_index) * ld_ctx + h * head_dim;
  for (int j = lane; j < head_dim; j += 32) {
    const float* v_col = base + 2 * hidden + j;
    float acc = 0.f;
    for (int key = 0; key < len; ++key) acc = fmaf(scores[key], v_col[static_cast<long long>(key) * ld], acc);
    out[j] = __float2bfloat16(acc * inv);
  }
}

}  // namespace aph

using namespace aph;

static unsigned blocks_for(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = 148ll * 16;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int aph_layernorm_any(const float* x, int64_t ld_x, int64_t rows, int32_t cols, const float* gamma, const float* beta, float eps,
                                 float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, void* stream_) {
  APH_REQUIRE(x && rows >= 0 && cols > 0, "layernorm_any: bad arguments");
  APH_REQUIRE((gamma == nullptr) == (beta == nullptr), "layernorm_any: gamma and beta come together");
  APH_REQUIRE(out_f32 || out_bf16, "layernorm_any: no output");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  layernorm_any_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, ld_x, rows, cols, gamma, beta, eps, out_f32, ld_f32,
                                                                                    static_cast<__nv_bfloat16*>(out_bf16), ld_bf16);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_add_sinusoidal(float* x, int64_t ld, int32_t n_utt, int32_t seq, int32_t cols, const float* bases, void* stream_) {
  APH_REQUIRE(x && bases && n_utt >= 0 && seq > 0 && cols > 0, "add_sinusoidal: bad arguments");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long rows = static_cast<long long>(n_utt) * seq;
  add_sinusoidal_kernel<<<blocks_for(rows * cols), 256, 0, stream>>>(x, ld, seq, cols, rows, bases);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_transpose_nfl(const float* in, int32_t n_utt, int32_t features, int32_t length, float* out, int64_t ld_out, void* stream_) {
  APH_REQUIRE(in && out && n_utt >= 0 && features > 0 && length > 0 && ld_out >= features, "transpose_nfl: bad arguments");
  APH_REQUIRE(n_utt <= 65535, "transpose_nfl: at most 65535 utterances");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid(static_cast<unsigned>((length + 31) / 32), static_cast<unsigned>((features + 31) / 32), static_cast<unsigned>(n_utt));
  transpose_nfl_kernel<<<grid, 256, 0, stream>>>(in, features, length, out, ld_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_reflect_pad_bf16(const float* x, int64_t ld_x, const int32_t* lengths, int32_t n_utt, int32_t length, int32_t channels,
                                    int32_t left, int32_t right, int32_t reflect, void* out_bf16, void* stream_) {
  APH_REQUIRE(x && lengths && out_bf16 && n_utt >= 0 && length > 0 && channels > 0 && left >= 0 && right >= 0, "reflect_pad: bad arguments");
  APH_REQUIRE(n_utt <= 65535, "reflect_pad: at most 65535 utterances");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long per_utt = static_cast<long long>(length + left + right) * channels;
  dim3 grid(blocks_for(per_utt), static_cast<unsigned>(n_utt));
  reflect_pad_kernel<<<grid, 256, 0, stream>>>(x, ld_x, lengths, length, channels, left, right, reflect, static_cast<__nv_bfloat16*>(out_bf16));
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_glu_rows(const float* y, int64_t ld_y, int64_t rows, int32_t out_channels, float* out, int64_t ld_out, void* stream_) {
  APH_REQUIRE(y && out && rows >= 0 && out_channels > 0 && ld_y >= 2 * out_channels && ld_out >= out_channels, "glu_rows: bad arguments");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  glu_rows_kernel<<<blocks_for(rows * out_channels), 256, 0, stream>>>(y, ld_y, rows, out_channels, out, ld_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_layernorm_any_backward(const float* x, int64_t ld_x, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols,
                                          const float* gamma, float eps, const float* resid, int64_t ld_resid, float* dx, int64_t ld_dx,
                                          float* dgamma, float* dbeta, void* stream_) {
  APH_REQUIRE(x && dy && rows >= 0 && cols > 0, "layernorm_any_backward: bad arguments");
  APH_REQUIRE(dx || dgamma, "layernorm_any_backward: no output");
  APH_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_any_backward: dgamma and dbeta come together");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  layernorm_any_backward_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, resid, ld_resid,
                                                                                             dx, ld_dx, dgamma, dbeta);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_activation_backward(float* d, int64_t ld_d, const float* y, int64_t ld_y, int64_t rows, int32_t cols, int32_t kind,
                                       void* out_bf16, int64_t ld_bf16, void* stream_) {
  APH_REQUIRE(d && y && rows >= 0 && cols > 0 && (kind == 2 || kind == 3), "activation_backward: kind 2 (ReLU) or 3 (LeakyReLU)");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  activation_backward_kernel<<<blocks_for(rows * cols), 256, 0, stream>>>(d, ld_d, y, ld_y, rows, cols, kind, static_cast<__nv_bfloat16*>(out_bf16),
                                                                          ld_bf16);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_glu_backward_bf16(const float* y, int64_t ld_y, const float* d_out, int64_t ld_d, int64_t rows, int32_t out_channels,
                                     void* dy_bf16, int64_t ld_dy, void* stream_) {
  APH_REQUIRE(y && d_out && dy_bf16 && rows >= 0 && out_channels > 0 && ld_y >= 2 * out_channels && ld_dy >= 2 * out_channels, "glu_backward: bad arguments");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  glu_backward_kernel<<<blocks_for(rows * out_channels), 256, 0, stream>>>(y, ld_y, d_out, ld_d, rows, out_channels, static_cast<__nv_bfloat16*>(dy_bf16),
                                                                            ld_dy);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_conv_input_backward(const float* d_cols, const int32_t* lengths, int32_t n_utt, int32_t length, int32_t channels,
                                       int32_t out_len, int32_t kernel, int32_t stride, int32_t left, int32_t right, int32_t reflect,
                                       float* d_x, int64_t ld_dx, void* stream_) {
  APH_REQUIRE(d_cols && lengths && d_x && n_utt >= 0 && length > 0 && channels > 0 && out_len > 0 && kernel > 0 && stride > 0, "conv_input_backward: bad arguments");
  APH_REQUIRE(n_utt <= 65535 && ld_dx >= channels, "conv_input_backward: at most 65535 utterances");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid(blocks_for(static_cast<long long>(length) * channels), static_cast<unsigned>(n_utt));
  conv_input_backward_kernel<<<grid, 256, 0, stream>>>(d_cols, lengths, n_utt, length, channels, out_len, kernel, stride, left, right, reflect, d_x, ld_dx);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_attention_small(const float* qkv, int64_t ld, void* ctx_bf16, int64_t ld_ctx, const int32_t* lengths, int32_t n_utt,
                                   int32_t heads, int32_t T, int32_t head_dim, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(qkv && ctx_bf16 && lengths, "null pointer");
  APH_REQUIRE(n_utt > 0 && heads > 0 && T > 0 && head_dim > 0, "empty problem");
  APH_REQUIRE(ld >= 3LL * heads * head_dim && ld_ctx >= static_cast<int64_t>(heads) * head_dim, "leading dimensions too small");
  const size_t smem = sizeof(float) * (static_cast<size_t>(kSmallAttWarps) * T + static_cast<size_t>(kSmallAttWarps) * head_dim);
  APH_REQUIRE(smem <= 200 * 1024, "sequence too long for the shared-memory staged scores");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  const dim3 grid(ceil_div(T, kSmallAttWarps), static_cast<unsigned>(n_utt) * heads);
  attention_small_kernel<<<grid, kSmallAttWarps * 32, smem, stream>>>(qkv, ld, static_cast<__nv_bfloat16*>(ctx_bf16), ld_ctx, lengths, T, heads,
                                                                      head_dim, 1.0f / sqrtf(static_cast<float>(head_dim)));
  APH_POST_LAUNCH(1);
  return APH_OK;
}
