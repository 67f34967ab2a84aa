// allophant_b200 — training-side kernels of the classifier heads.
//
//   aph_allophone_forward / _backward   AllophoneMapping.map_allophones + _multiply_allophone_matrix
//                                       (acoustic_model.py:75-87, 142-159) and its autograd backward
//   aph_transpose_cast_bf16             fp32/bf16 [R][C] -> bf16 [C][R_pad]: makes the batch axis the
//                                       contiguous K axis of the weight-gradient GEMMs (dW = dY^T X)
//   aph_colsum_f32                      bias gradients (column sums of dY)
//   aph_embedding_bag_backward          EmbeddingBag(mode="sum") backward of the composition layer
//                                       (acoustic_model.py:208, 225-232)
//   aph_softmax_backward_cols           backward of softmax(dependency logits) (acoustic_model.py:497-514)
//
// The reference materialises [T', P, Q] per utterance for the allophone layer; here the fixed
// sparsity of the allophone mask (a handful of phones per phoneme) is walked through a CSR list.
#include "aph_common.cuh"

namespace aph {

constexpr float kPadValue = -3.4028234663852886e+38f;  // torch.finfo(torch.float32).min, acoustic_model.py:72

// out[n][t][q] = max over phones p listed for (language(n), q) of logits[n][t][p] * W[lang][p][q];
// phonemes without any phone keep the pad value.  arg[n][t][q] = winning p (or -1).
__global__ void __launch_bounds__(256) allophone_forward_kernel(const float* __restrict__ logits, long long stride_n,
                                                                long long stride_t, int n_utt, int T, int P1, int Q1,
                                                                const float* __restrict__ W, const int* __restrict__ csr_off,
                                                                const int* __restrict__ csr_p,
                                                                const long long* __restrict__ language_ids,
                                                                float* __restrict__ out, int* __restrict__ arg) {
  const long long total = static_cast<long long>(n_utt) * T * Q1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % Q1);
    const long long nt = i / Q1;
    const int t = static_cast<int>(nt % T);
    const int n = static_cast<int>(nt / T);
    const int lang = static_cast<int>(language_ids[n]);
    const float* row = logits + n * stride_n + t * stride_t;
    const float* w = W + static_cast<long long>(lang) * P1 * Q1 + q;
    const int lo = csr_off[lang * Q1 + q], hi = csr_off[lang * Q1 + q + 1];
    float best = kPadValue;
    int best_p = -1;
    for (int k = lo; k < hi; ++k) {
      const int p = csr_p[k];
      const float v = row[p] * w[static_cast<long long>(p) * Q1];
      if (v > best || best_p < 0) {
        best = v;
        best_p = p;
      }
    }
    out[i] = best;
    if (arg) arg[i] = best_p;
  }
}

// grad_logits[n][t][p*] += g * W[lang][p*][q];  grad_W[lang][p*][q] += g * logits[n][t][p*]
__global__ void __launch_bounds__(256) allophone_backward_kernel(const float* __restrict__ grad_out, const int* __restrict__ arg,
                                                                 const float* __restrict__ logits, long long stride_n,
                                                                 long long stride_t, int n_utt, int T, int P1, int Q1,
                                                                 const float* __restrict__ W,
                                                                 const long long* __restrict__ language_ids,
                                                                 float* __restrict__ grad_logits /*[n][t][P1] contiguous*/,
                                                                 float* __restrict__ grad_W) {
  const long long total = static_cast<long long>(n_utt) * T * Q1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = arg[i];
    const float g = grad_out[i];
    if (p < 0 || g == 0.f) continue;
    const int q = static_cast<int>(i % Q1);
    const long long nt = i / Q1;
    const int t = static_cast<int>(nt % T);
    const int n = static_cast<int>(nt / T);
    const int lang = static_cast<int>(language_ids[n]);
    const long long widx = (static_cast<long long>(lang) * P1 + p) * Q1 + q;
    if (grad_logits) atomicAdd(grad_logits + nt * P1 + p, g * W[widx]);
    if (grad_W) atomicAdd(grad_W + widx, g * logits[n * stride_n + t * stride_t + p]);
  }
}

// 32x32 shared-memory tile transpose with conversion to bf16; rows beyond R are zero in the padded output
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const TIn* __restrict__ in, long long ld_in, long long R, int C,
                                                             __nv_bfloat16* __restrict__ out, long long ld_out, long long R_pad) {
  __shared__ float tile[32][33];
  const long long r0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long r = r0 + ty + 8 * k;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      if constexpr (sizeof(TIn) == 2) {
        v = __bfloat162float(in[r * ld_in + c]);
      } else {
        v = in[r * ld_in + c];
      }
    }
    tile[ty + 8 * k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long r = r0 + tx;
    if (c < C && r < R_pad) out[static_cast<long long>(c) * ld_out + r] = __float2bfloat16(tile[tx][ty + 8 * k]);
  }
}

template <typename TIn>
__global__ void __launch_bounds__(256) colsum_kernel(const TIn* __restrict__ in, long long ld, long long R, int C,
                                                     float* __restrict__ out) {
  // block = 32 x VEC columns x 8 row-lanes, 16-byte loads (VEC = 8 bf16 / 4 fp32 per thread); grid.y splits the rows;
  // partial sums combined through smem, then one atomic per column and block (out pre-zeroed)
  constexpr int VEC = sizeof(TIn) == 2 ? 8 : 4;
  __shared__ float part[8][32 * VEC + 1];
  const int lane = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  const long long rows_per = (R + gridDim.y - 1) / gridDim.y;
  const long long lo = blockIdx.y * rows_per, hi = min(R, lo + rows_per);
  float s[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) s[v] = 0.f;
  const bool vector_ok = c0 + VEC <= C && (ld % VEC) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  if (vector_ok) {
#pragma unroll 4
    for (long long r = lo + ty; r < hi; r += 8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(in + r * ld + c0);
      if constexpr (sizeof(TIn) == 2) {
        const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
        s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
      } else {
        s[0] += __uint_as_float(raw.x); s[1] += __uint_as_float(raw.y); s[2] += __uint_as_float(raw.z); s[3] += __uint_as_float(raw.w);
      }
    }
  } else if (c0 < C) {
    for (long long r = lo + ty; r < hi; r += 8)
      for (int v = 0; v < VEC && c0 + v < C; ++v) {
        if constexpr (sizeof(TIn) == 2) {
          s[v] += __bfloat162float(in[r * ld + c0 + v]);
        } else {
          s[v] += in[r * ld + c0 + v];
        }
      }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) part[ty][lane * VEC + v] = s[v];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VEC; i += 256) {
    const int c = blockIdx.x * 32 * VEC + i;
    if (c < C) {
      float total = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) total += part[k][i];
      atomicAdd(out + c, total);
    }
  }
}

// grad_weight[idx] += grad_rows[row]: row 0 -> category 0 (blank); row 1+v -> tfi[v][f] + offsets[f] for every f
__global__ void __launch_bounds__(128) embedding_bag_backward_kernel(const float* __restrict__ grad_rows, long long ld, int V,
                                                                     int F, int E, const long long* __restrict__ tfi,
                                                                     const long long* __restrict__ offsets,
                                                                     float* __restrict__ grad_weight) {
  const int row = blockIdx.x;  // 0 .. V
  const float* g = grad_rows + static_cast<long long>(row) * ld;
  if (row == 0) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(grad_weight + e, g[e]);
    return;
  }
  const long long* idx = tfi + static_cast<long long>(row - 1) * F;
  for (int f = 0; f < F; ++f) {
    const long long cat = idx[f] + (offsets ? offsets[f] : 0);
    float* dst = grad_weight + cat * E;
    for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(dst + e, g[e]);
  }
}

// d logits[skip:] += p * (dp - sum(p * dp)) for each dependency block, p = bf16 probabilities stored in X
__global__ void __launch_bounds__(256) softmax_backward_cols_kernel(const float* __restrict__ grad_x, long long ld_gx,
                                                                    const __nv_bfloat16* __restrict__ x, long long ld_x,
                                                                    long long rows, const int* __restrict__ x_col,
                                                                    const int* __restrict__ width, const int* __restrict__ dst_col,
                                                                    int n_deps, int skip, float* __restrict__ grad_logits,
                                                                    long long ld_gl) {
  const long long total = rows * n_deps;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / n_deps;
    const int d = static_cast<int>(i - row * n_deps);
    const int w = width[d] - skip;
    const float* gp = grad_x + row * ld_gx + x_col[d];
    const __nv_bfloat16* p = x + row * ld_x + x_col[d];
    float dot = 0.f;
    for (int c = 0; c < w; ++c) dot += __bfloat162float(p[c]) * gp[c];
    float* dst = grad_logits + row * ld_gl + dst_col[d] + skip;
    for (int c = 0; c < w; ++c) dst[c] += __bfloat162float(p[c]) * (gp[c] - dot);
  }
}


// ------------------------------------------------------------------------------------------------
// LayerNorm backward (autograd of nn.LayerNorm, HF:429-431, 766-767, 792): one warp per row,
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += dy * xhat;  dbeta += dy
// Statistics are recomputed from x (two-pass, fp32) exactly like the forward kernel, so nothing but
// the layer input has to be kept.  dx is added to `dx_resid` (the gradient already flowing through
// the residual connection) when given.  dgamma/dbeta: per-thread partials -> block reduction in
// shared memory -> one atomicAdd per column and block (outputs pre-zeroed by the launcher).
template <typename TX, typename TDY, int COLS>
__global__ void __launch_bounds__(256) layernorm_backward_kernel(const TX* __restrict__ x, long long ld_x,
                                                                 const TDY* __restrict__ dy, long long ld_dy, long long rows,
                                                                 const float* __restrict__ gamma, float eps,
                                                                 const float* __restrict__ dx_resid, long long ld_resid,
                                                                 float* __restrict__ dx, long long ld_dx,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int V = COLS / 128;  // float4 groups per lane
  __shared__ float red[8][COLS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float gm[V][4];
  float acc_g[V][4], acc_b[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + 4 * (lane + 32 * i));
    gm[i][0] = g4.x; gm[i][1] = g4.y; gm[i][2] = g4.z; gm[i][3] = g4.w;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc_g[i][e] = acc_b[i][e] = 0.f;
  }
  const float inv_n = 1.0f / COLS;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows; row += static_cast<long long>(gridDim.x) * 8) {
    float xv[V][4], gv[V][4];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = 4 * (lane + 32 * i);
      if constexpr (sizeof(TX) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(x + row * ld_x + c);
        xv[i][0] = t.x; xv[i][1] = t.y; xv[i][2] = t.z; xv[i][3] = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(x + row * ld_x + c);
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        xv[i][0] = a.x; xv[i][1] = a.y; xv[i][2] = b.x; xv[i][3] = b.y;
      }
      if constexpr (sizeof(TDY) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(dy + row * ld_dy + c);
        gv[i][0] = t.x; gv[i][1] = t.y; gv[i][2] = t.z; gv[i][3] = t.w;
      } else {
        const uint2 t = *reinterpret_cast<const uint2*>(dy + row * ld_dy + c);
        const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
        gv[i][0] = a.x; gv[i][1] = a.y; gv[i][2] = b.x; gv[i][3] = b.y;
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s += xv[i][e];
    const float mean = warp_sum(s) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = xv[i][e] - mean;
        sq = fmaf(d, d, sq);
      }
    const float rstd = rsqrtf(warp_sum(sq) * inv_n + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (xv[i][e] - mean) * rstd;
        const float d = gv[i][e];
        acc_g[i][e] = fmaf(d, xh, acc_g[i][e]);
        acc_b[i][e] += d;
        const float g = d * gm[i][e];
        xv[i][e] = xh;
        gv[i][e] = g;
        s1 += g;
        s2 = fmaf(g, xh, s2);
      }
    s1 = warp_sum(s1) * inv_n;
    s2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = 4 * (lane + 32 * i);
      float4 o;
      o.x = rstd * (gv[i][0] - s1 - xv[i][0] * s2);
      o.y = rstd * (gv[i][1] - s1 - xv[i][1] * s2);
      o.z = rstd * (gv[i][2] - s1 - xv[i][2] * s2);
      o.w = rstd * (gv[i][3] - s1 - xv[i][3] * s2);
      if (dx_resid != nullptr) {
        const float4 r4 = *reinterpret_cast<const float4*>(dx_resid + row * ld_resid + c);
        o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
      }
      *reinterpret_cast<float4*>(dx + row * ld_dx + c) = o;
    }
  }
  if (dgamma == nullptr && dbeta == nullptr) return;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma : dbeta;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) red[warp][4 * (lane + 32 * i) + e] = pass == 0 ? acc_g[i][e] : acc_b[i][e];
    __syncthreads();
    if (dst != nullptr) {
      for (int c = threadIdx.x; c < COLS; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += red[w8][c];
        atomicAdd(dst + c, t);
      }
    }
  }
}

// rows of padded frames -> 0 (backward of `hidden_states[~mask] = 0`, HF:753-756)
__global__ void __launch_bounds__(256) mask_rows_kernel(float* __restrict__ x, long long ld, long long rows, int cols,
                                                        const int* __restrict__ lengths, int period) {
  const int vec = cols >> 2;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec;
    const int c = static_cast<int>(i - row * vec) * 4;
    const long long utt = row / period;
    const int t = static_cast<int>(row - utt * period);
    if (t >= lengths[utt]) *reinterpret_cast<float4*>(x + row * ld + c) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// dst[r][c] += src[r][c]
__global__ void __launch_bounds__(256) add_2d_kernel(float* __restrict__ dst, long long ld_dst, const float* __restrict__ src,
                                                     long long ld_src, long long rows, int cols) {
  const int vec = cols >> 2;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec;
    const int c = static_cast<int>(i - row * vec) * 4;
    float4 a = *reinterpret_cast<float4*>(dst + row * ld_dst + c);
    const float4 b = *reinterpret_cast<const float4*>(src + row * ld_src + c);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    *reinterpret_cast<float4*>(dst + row * ld_dst + c) = a;
  }
}

// y = dy * gelu'(pre) as bf16 (input of the positional conv's dgrad / wgrad GEMMs)
__global__ void __launch_bounds__(256) gelu_backward_kernel(const float* __restrict__ dy, long long ld_dy,
                                                            const __nv_bfloat16* __restrict__ pre, long long ld_pre,
                                                            long long rows, int cols, __nv_bfloat16* __restrict__ out,
                                                            long long ld_out) {
  const int vec = cols >> 2;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec;
    const int c = static_cast<int>(i - row * vec) * 4;
    const float4 d = *reinterpret_cast<const float4*>(dy + row * ld_dy + c);
    const uint2 pw = *reinterpret_cast<const uint2*>(pre + row * ld_pre + c);
    const float2 p0 = unpack_bf16x2(pw.x), p1 = unpack_bf16x2(pw.y);
    uint2 o;
    o.x = pack_bf16x2(d.x * gelu_erf_grad(p0.x), d.y * gelu_erf_grad(p0.y));
    o.y = pack_bf16x2(d.z * gelu_erf_grad(p1.x), d.w * gelu_erf_grad(p1.y));
    *reinterpret_cast<uint2*>(out + row * ld_out + c) = o;
  }
}

// ---- positional conv (weight_norm(dim=2) grouped Conv1d, HF:326-350): backward-side packing ----
// dgrad B operand for the sliding-tap GEMM: dst[g*Cg + ci][j'*Cg + co] = w[g*Cg + co][ci][K-1-j'],
// w = g[tap] * v / ||v[:,:,tap]||  (dx[s] = sum_j dy[s + K/2 - j] w[j]: taps flipped, channels swapped)
__global__ void __launch_bounds__(256) pack_posconv_dgrad_kernel(const float* __restrict__ weight_g, const float* __restrict__ v,
                                                                 const float* __restrict__ tap_scale /*[K] g/norm*/,
                                                                 __nv_bfloat16* __restrict__ dst, int O, int Cg, int K) {
  const long long total = static_cast<long long>(O) * K * Cg;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(i % Cg);
    const long long rest = i / Cg;
    const int jp = static_cast<int>(rest % K);
    const int n = static_cast<int>(rest / K);  // g*Cg + ci
    const int g = n / Cg, ci = n - g * Cg;
    const int tap = K - 1 - jp;
    dst[i] = __float2bfloat16(v[(static_cast<long long>(g * Cg + co) * Cg + ci) * K + tap] * tap_scale[tap]);
  }
}

// per tap: norm^2 = sum v^2, scale = g / norm
__global__ void __launch_bounds__(256) posconv_tap_norm_kernel(const float* __restrict__ weight_g, const float* __restrict__ v,
                                                               int O, int Cg, int K, float* __restrict__ tap_scale,
                                                               float* __restrict__ tap_norm2) {
  __shared__ float red[256];
  const int tap = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < O * Cg; i += 256) {
    const float x = v[static_cast<long long>(i) * K + tap];
    s = fmaf(x, x, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tap_norm2[tap] = red[0];
    tap_scale[tap] = weight_g[tap] * rsqrtf(red[0]);
  }
}

// raw[tap][o][256] (output of the DIAG_TAPS GEMM) -> dot[tap] = sum_{o,ci} dW[o][ci][tap] * v[o][ci][tap]
__global__ void __launch_bounds__(256) posconv_wgrad_dot_kernel(const float* __restrict__ raw, const float* __restrict__ v, int O,
                                                                int Cg, int K, int W, float* __restrict__ dot) {
  __shared__ float red[256];
  const int tap = blockIdx.x;
  const int blocks = W / Cg;  // channel groups per W-wide diagonal block (256: DIAG_TAPS GEMM; O: one full matrix per tap)
  float s = 0.f;
  for (int i = threadIdx.x; i < O * Cg; i += 256) {
    const int o = i / Cg, ci = i - o * Cg;
    const float dw = raw[(static_cast<long long>(tap) * O + o) * W + ((o / Cg) % blocks) * Cg + ci];
    s = fmaf(dw, v[static_cast<long long>(i) * K + tap], s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dot[tap] = red[0];
}

// dv = g/norm * (dW - v * dot / norm^2);  dg[tap] = dot / norm
__global__ void __launch_bounds__(256) posconv_wgrad_finish_kernel(const float* __restrict__ raw, const float* __restrict__ v,
                                                                   const float* __restrict__ weight_g, const float* __restrict__ dot,
                                                                   const float* __restrict__ tap_norm2, int O, int Cg, int K, int W,
                                                                   float* __restrict__ grad_v, float* __restrict__ grad_g) {
  const int blocks = W / Cg;
  const long long total = static_cast<long long>(O) * Cg * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % K);
    const long long oc = i / K;
    const int o = static_cast<int>(oc / Cg), ci = static_cast<int>(oc - static_cast<long long>(o) * Cg);
    const float n2 = tap_norm2[tap];
    const float inv = rsqrtf(n2);
    const float dw = raw[(static_cast<long long>(tap) * O + o) * W + ((o / Cg) % blocks) * Cg + ci];
    grad_v[i] = weight_g[tap] * inv * (dw - v[i] * dot[tap] / n2);
    if (oc == 0) grad_g[tap] = dot[tap] * inv;
  }
}

static inline unsigned blocks_for(long long items, int per_block, long long cap) {
  long long b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<unsigned>(b);
}

}  // namespace aph

using namespace aph;

extern "C" int aph_allophone_forward(const float* logits, int64_t stride_n, int64_t stride_t, int32_t n_utt, int32_t T,
                                     int32_t n_phones, int32_t n_phonemes, const float* matrices, const int32_t* csr_offsets,
                                     const int32_t* csr_phones, const int64_t* language_ids, float* out, int32_t* argmax_out,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(logits && matrices && csr_offsets && csr_phones && language_ids && out, "null pointer");
  APH_REQUIRE(n_utt > 0 && T > 0 && n_phones > 0 && n_phonemes > 0, "empty problem");
  const long long total = static_cast<long long>(n_utt) * T * n_phonemes;
  allophone_forward_kernel<<<blocks_for(total, 256, 16LL * sm_count()), 256, 0, stream>>>(
      logits, stride_n, stride_t, n_utt, T, n_phones, n_phonemes, matrices, csr_offsets, csr_phones,
      reinterpret_cast<const long long*>(language_ids), out, argmax_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_allophone_backward(const float* grad_out, const int32_t* argmax_in, const float* logits, int64_t stride_n,
                                      int64_t stride_t, int32_t n_utt, int32_t T, int32_t n_phones, int32_t n_phonemes,
                                      const float* matrices, const int64_t* language_ids, float* grad_logits, float* grad_matrices,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(grad_out && argmax_in && logits && matrices && language_ids && (grad_logits || grad_matrices), "null pointer");
  APH_REQUIRE(n_utt > 0 && T > 0 && n_phones > 0 && n_phonemes > 0, "empty problem");
  const long long total = static_cast<long long>(n_utt) * T * n_phonemes;
  allophone_backward_kernel<<<blocks_for(total, 256, 16LL * sm_count()), 256, 0, stream>>>(
      grad_out, argmax_in, logits, stride_n, stride_t, n_utt, T, n_phones, n_phonemes, matrices,
      reinterpret_cast<const long long*>(language_ids), grad_logits, grad_matrices);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_transpose_cast_bf16(const void* in, int32_t in_is_f32, int64_t ld_in, int64_t rows, int32_t cols, void* out_bf16,
                                       int64_t ld_out, int64_t rows_padded, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(in && out_bf16, "null pointer");
  APH_REQUIRE(rows >= 0 && cols > 0 && rows_padded >= rows && ld_out >= rows_padded, "bad shape");
  if (rows_padded == 0) return APH_OK;
  dim3 grid(static_cast<unsigned>((rows_padded + 31) / 32), static_cast<unsigned>((cols + 31) / 32));
  if (in_is_f32)
    transpose_cast_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), ld_in, rows, cols,
                                                           static_cast<__nv_bfloat16*>(out_bf16), ld_out, rows_padded);
  else
    transpose_cast_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld_in, rows, cols,
                                                                   static_cast<__nv_bfloat16*>(out_bf16), ld_out, rows_padded);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

static int colsum_any(const void* in, bool is_f32, int64_t ld, int64_t rows, int32_t cols, float* out, cudaStream_t stream) {
  APH_REQUIRE(in && out, "null pointer");
  APH_REQUIRE(cols > 0, "bad shape");
  APH_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * cols, stream));
  if (rows <= 0) return APH_OK;
  const int per_block = is_f32 ? 128 : 256;  // columns per block (32 lanes x 16 bytes)
  const unsigned gx = static_cast<unsigned>((cols + per_block - 1) / per_block);
  // ~4 blocks per SM in total, at least 64 rows per block
  unsigned gy = static_cast<unsigned>(std::min<long long>((rows + 63) / 64, std::max<long long>(1, (148 * 4 + gx - 1) / gx)));
  if (gy < 1) gy = 1;
  const dim3 grid(gx, gy);
  if (is_f32)
    colsum_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), ld, rows, cols, out);
  else
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), ld, rows, cols, out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_colsum_f32(const float* in, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream_) {
  return colsum_any(in, true, ld, rows, cols, out, static_cast<cudaStream_t>(stream_));
}

extern "C" int aph_colsum_bf16(const void* in, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream_) {
  return colsum_any(in, false, ld, rows, cols, out, static_cast<cudaStream_t>(stream_));
}

extern "C" int aph_embedding_bag_backward(const float* grad_rows, int64_t ld, int32_t n_phonemes, int32_t n_features,
                                          int32_t embedding_size, const int64_t* tfi, const int64_t* category_offsets,
                                          float* grad_weight, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(grad_rows && tfi && grad_weight, "null pointer");
  APH_REQUIRE(n_phonemes >= 0 && n_features > 0 && embedding_size > 0, "bad shape");
  embedding_bag_backward_kernel<<<n_phonemes + 1, 128, 0, stream>>>(grad_rows, ld, n_phonemes, n_features, embedding_size,
                                                                    reinterpret_cast<const long long*>(tfi),
                                                                    reinterpret_cast<const long long*>(category_offsets),
                                                                    grad_weight);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_softmax_backward_cols(const float* grad_x, int64_t ld_gx, const void* x_bf16, int64_t ld_x, int64_t rows,
                                         const int32_t* x_col, const int32_t* width, const int32_t* dst_col, int32_t n_deps,
                                         int32_t skip, float* grad_logits, int64_t ld_gl, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(grad_x && x_bf16 && x_col && width && dst_col && grad_logits, "null pointer");
  if (rows <= 0 || n_deps <= 0) return APH_OK;
  softmax_backward_cols_kernel<<<blocks_for(rows * n_deps, 256, 8LL * sm_count()), 256, 0, stream>>>(
      grad_x, ld_gx, static_cast<const __nv_bfloat16*>(x_bf16), ld_x, rows, x_col, width, dst_col, n_deps, skip, grad_logits, ld_gl);
  APH_POST_LAUNCH(1);
  return APH_OK;
}


template <typename TX, typename TDY>
static int launch_ln_backward(const void* x, int64_t ld_x, const void* dy, int64_t ld_dy, int64_t rows, int32_t cols,
                              const float* gamma, float eps, const float* dx_resid, int64_t ld_resid, float* dx, int64_t ld_dx,
                              float* dgamma, float* dbeta, cudaStream_t stream) {
  long long blocks = (rows + 7) / 8;
  const long long cap = 2LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (cols == 1024)
    layernorm_backward_kernel<TX, TDY, 1024><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const TX*>(x), ld_x, static_cast<const TDY*>(dy), ld_dy, rows, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta);
  else
    layernorm_backward_kernel<TX, TDY, 512><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const TX*>(x), ld_x, static_cast<const TDY*>(dy), ld_dy, rows, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_layernorm_backward(const void* x, int32_t x_is_f32, int64_t ld_x, const void* dy, int32_t dy_is_f32,
                                      int64_t ld_dy, int64_t rows, int32_t cols, const float* gamma, float eps,
                                      const float* dx_resid, int64_t ld_resid, float* dx, int64_t ld_dx, float* dgamma,
                                      float* dbeta, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && dy && gamma && dx, "null pointer");
  APH_REQUIRE(cols == 512 || cols == 1024, "LayerNorm backward supports 512 and 1024 columns");
  APH_REQUIRE(ld_x % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0 && ld_resid % 4 == 0, "leading dimensions must be multiples of 4");
  if (dgamma) APH_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, stream));
  if (dbeta) APH_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, stream));
  if (rows <= 0) return APH_OK;
  if (x_is_f32 && dy_is_f32)
    return launch_ln_backward<float, float>(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta, stream);
  if (x_is_f32)
    return launch_ln_backward<float, __nv_bfloat16>(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta, stream);
  if (dy_is_f32)
    return launch_ln_backward<__nv_bfloat16, float>(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta, stream);
  return launch_ln_backward<__nv_bfloat16, __nv_bfloat16>(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta, stream);
}

extern "C" int aph_mask_rows_f32(float* x, int64_t ld, int64_t rows, int32_t cols, const int32_t* lengths, int32_t len_period,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && lengths, "null pointer");
  APH_REQUIRE(cols > 0 && cols % 4 == 0 && ld % 4 == 0 && len_period > 0, "bad shape");
  if (rows <= 0) return APH_OK;
  mask_rows_kernel<<<blocks_for(rows * (cols / 4), 256, 16LL * sm_count()), 256, 0, stream>>>(x, ld, rows, cols, lengths, len_period);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_gelu_backward_bf16(const float* dy, int64_t ld_dy, const void* pre_bf16, int64_t ld_pre, int64_t rows,
                                      int32_t cols, void* out_bf16, int64_t ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(dy && pre_bf16 && out_bf16, "null pointer");
  APH_REQUIRE(cols > 0 && cols % 4 == 0 && ld_dy % 4 == 0 && ld_pre % 4 == 0 && ld_out % 4 == 0, "bad shape");
  if (rows <= 0) return APH_OK;
  gelu_backward_kernel<<<blocks_for(rows * (cols / 4), 256, 16LL * sm_count()), 256, 0, stream>>>(
      dy, ld_dy, static_cast<const __nv_bfloat16*>(pre_bf16), ld_pre, rows, cols, static_cast<__nv_bfloat16*>(out_bf16), ld_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_pack_posconv_weight_dgrad(const float* weight_g, const float* weight_v, void* dst_bf16, float* tap_scratch,
                                             int32_t out_channels, int32_t group_channels, int32_t kernel, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(weight_g && weight_v && dst_bf16 && tap_scratch, "null pointer");
  APH_REQUIRE(out_channels > 0 && group_channels > 0 && kernel > 0 && out_channels % group_channels == 0, "bad shape");
  posconv_tap_norm_kernel<<<kernel, 256, 0, stream>>>(weight_g, weight_v, out_channels, group_channels, kernel, tap_scratch,
                                                      tap_scratch + kernel);
  const long long total = static_cast<long long>(out_channels) * kernel * group_channels;
  pack_posconv_dgrad_kernel<<<blocks_for(total, 256, 16LL * sm_count()), 256, 0, stream>>>(
      weight_g, weight_v, tap_scratch, static_cast<__nv_bfloat16*>(dst_bf16), out_channels, group_channels, kernel);
  APH_POST_LAUNCH(2);
  return APH_OK;
}

extern "C" int aph_posconv_weight_backward_blocks(const float* raw, const float* weight_g, const float* weight_v, float* tap_scratch,
                                                  int32_t out_channels, int32_t group_channels, int32_t kernel, int32_t block_width,
                                                  float* grad_g, float* grad_v, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(raw && weight_g && weight_v && tap_scratch && grad_g && grad_v, "null pointer");
  APH_REQUIRE(out_channels > 0 && group_channels > 0 && kernel > 0 && block_width > 0 && block_width % group_channels == 0 &&
                  out_channels % block_width == 0,
              "bad shape: the diagonal blocks must hold whole channel groups and tile the channels");
  float* scale = tap_scratch;
  float* norm2 = tap_scratch + kernel;
  float* dot = tap_scratch + 2 * kernel;
  posconv_tap_norm_kernel<<<kernel, 256, 0, stream>>>(weight_g, weight_v, out_channels, group_channels, kernel, scale, norm2);
  posconv_wgrad_dot_kernel<<<kernel, 256, 0, stream>>>(raw, weight_v, out_channels, group_channels, kernel, block_width, dot);
  const long long total = static_cast<long long>(out_channels) * group_channels * kernel;
  posconv_wgrad_finish_kernel<<<blocks_for(total, 256, 16LL * sm_count()), 256, 0, stream>>>(
      raw, weight_v, weight_g, dot, norm2, out_channels, group_channels, kernel, block_width, grad_v, grad_g);
  APH_POST_LAUNCH(3);
  return APH_OK;
}

extern "C" int aph_posconv_weight_backward(const float* raw, const float* weight_g, const float* weight_v, float* tap_scratch,
                                           int32_t out_channels, int32_t group_channels, int32_t kernel, float* grad_g,
                                           float* grad_v, void* stream_) {
  return aph_posconv_weight_backward_blocks(raw, weight_g, weight_v, tap_scratch, out_channels, group_channels, kernel, 256, grad_g, grad_v,
                                            stream_);
}

extern "C" int aph_add_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int32_t cols,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(dst && src, "null pointer");
  APH_REQUIRE(cols > 0 && cols % 4 == 0 && ld_dst % 4 == 0 && ld_src % 4 == 0, "bad shape");
  if (rows <= 0) return APH_OK;
  add_2d_kernel<<<blocks_for(rows * (cols / 4), 256, 16LL * sm_count()), 256, 0, stream>>>(dst, ld_dst, src, ld_src, rows, cols);
  APH_POST_LAUNCH(1);
  return APH_OK;
}
