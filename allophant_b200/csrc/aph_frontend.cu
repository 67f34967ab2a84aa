// allophant_b200 — memory-bound front-end kernels of the acoustic encoder.
//
//   aph_wave_stats / aph_wave_norm   zero_mean_unit_var_norm   acoustic_model.py:762-767
//   aph_frame_lengths                conv_length chain          frontend.py:192-203, acoustic_model.py:832-835
//   aph_conv0_ln_gelu                Conv1d(1,512,10,5)+LayerNorm(512)+GELU   HF:275-299 (layer 0)
//   aph_conv0_gn_gelu (+stats)       Conv1d(1,512,10,5,no bias)+GroupNorm(512,512)+GELU  HF:302-323
//   aph_layernorm_rows               LayerNorm(+GELU) over channels-last rows  HF:290-299, 429-431, 766-767, 792
//
// All of them are HBM-bound: every input element is read once and every output
// element written once, with 16-byte vector accesses along the contiguous axis.
#include "aph_common.cuh"

namespace aph {

// ---------------------------------------------------------------------------
// waveform statistics: per utterance  S0 = sum_all x, S1 = sum_valid x, S2 = sum_valid x^2
// (the reference divides the sum over the WHOLE padded row by the length,
// acoustic_model.py:764, and masks the deviations afterwards, 765-766)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wave_stats_kernel(const float* __restrict__ x, const long long* __restrict__ lengths,
                                                         int T, double* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long len = lengths[b];
  const float* row = x + static_cast<long long>(b) * T;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const int chunk = (T + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * chunk;
  const int hi = min(T, lo + chunk);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float v = row[i];
    s0 += v;
    if (i < len) {
      s1 += v;
      s2 += static_cast<double>(v) * v;
    }
  }
  __shared__ double red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = s0;
    red[1][warp] = s1;
    red[2][warp] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0, d = 0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      c += red[1][w];
      d += red[2][w];
    }
    atomicAdd(&stats[b * 3 + 0], a);
    atomicAdd(&stats[b * 3 + 1], c);
    atomicAdd(&stats[b * 3 + 2], d);
  }
}

// mean / rstd per utterance from the three sums (fp64, then rounded to fp32)
__global__ void wave_finalize_kernel(const double* __restrict__ stats, const long long* __restrict__ lengths, int n,
                                     float2* __restrict__ mean_rstd) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const double len = static_cast<double>(lengths[b]);
  const double mean = stats[b * 3 + 0] / len;
  // sum_valid (x - mean)^2 = S2 - 2 mean S1 + len mean^2
  double var = (stats[b * 3 + 2] - 2.0 * mean * stats[b * 3 + 1] + len * mean * mean) / len;
  if (var < 0.0) var = 0.0;
  mean_rstd[b] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + 1e-7)));
}

__global__ void __launch_bounds__(256) wave_norm_kernel(const float* __restrict__ x, const long long* __restrict__ lengths,
                                                        const float2* __restrict__ mean_rstd, int T,
                                                        float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long len = lengths[b];
  const float2 mr = mean_rstd[b];
  const long long base = static_cast<long long>(b) * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    out[base + i] = i < len ? (x[base + i] - mr.x) * mr.y : 0.f;
  }
}

// frames = fold over layers of floor((L - k)/s) + 1
__global__ void frame_lengths_kernel(const long long* __restrict__ lengths, int n, const int* __restrict__ kernels,
                                     const int* __restrict__ strides, int n_layers, int* __restrict__ frames32,
                                     long long* __restrict__ frames64) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  long long L = lengths[b];
  for (int i = 0; i < n_layers; ++i) {
    // torch.div(..., rounding_mode="floor") semantics for negative numerators
    long long num = L - kernels[i];
    long long q = num / strides[i];
    if ((num % strides[i] != 0) && ((num < 0) != (strides[i] < 0))) --q;
    L = q + 1;
  }
  if (frames32) frames32[b] = static_cast<int>(L);
  if (frames64) frames64[b] = L;
}

// ---------------------------------------------------------------------------
// conv layer 0 (C_in = 1, k = 10, s = 5) fused with the waveform normalisation,
// LayerNorm over the 512 output channels and GELU; bf16 channels-last output.
// One warp owns kTT consecutive output frames; lane l owns channels
// {2l, 2l+1} + 64 i  (i = 0..7) so every store instruction writes 128
// contiguous bytes.
// ---------------------------------------------------------------------------
constexpr int kC0 = 512;
constexpr int kK0 = 10;
constexpr int kS0 = 5;
constexpr int kTT = 4;

template <bool kLayerNorm>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ x, const long long* __restrict__ lengths,
                                                    const float2* __restrict__ mean_rstd, int T, int L0,
                                                    const float* __restrict__ w /*[512][10]*/,
                                                    const float* __restrict__ bias /*[512] or null*/,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float eps, __nv_bfloat16* __restrict__ out /*[N][L0][512]*/,
                                                    float* __restrict__ raw_out /*GroupNorm path: fp32 conv output*/,
                                                    double* __restrict__ gn_stats /*[N][512][2] or null*/,
                                                    int skip_padded) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float w_s[kK0][kC0];
  __shared__ __align__(16) float b_s[kC0], g_s[kC0], be_s[kC0];
  for (int i = threadIdx.x; i < kC0 * kK0; i += blockDim.x) w_s[i % kK0][i / kK0] = w[i];
  for (int i = threadIdx.x; i < kC0; i += blockDim.x) {
    b_s[i] = bias ? bias[i] : 0.f;
    g_s[i] = gamma ? gamma[i] : 1.f;
    be_s[i] = beta ? beta[i] : 0.f;
  }
  __syncthreads();

  const int b = blockIdx.y;
  const long long len = lengths ? lengths[b] : T;
  const float2 mr = mean_rstd ? mean_rstd[b] : make_float2(0.f, 1.f);
  // frames of this utterance that depend only on valid samples
  long long l0_valid = len >= kK0 ? (len - kK0) / kS0 + 1 : 0;
  if (l0_valid > L0 || !skip_padded) l0_valid = L0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float* row = x + static_cast<long long>(b) * T;

  double gsum[16], gsq[16];
  if (!kLayerNorm) {
#pragma unroll
    for (int i = 0; i < 16; ++i) gsum[i] = gsq[i] = 0.0;
  }

  for (int t0 = (blockIdx.x * warps_per_block + warp) * kTT; t0 < (kLayerNorm ? l0_valid : L0);
       t0 += gridDim.x * warps_per_block * kTT) {
    // window of (kTT-1)*5 + 10 = 25 samples, one per lane
    const int xi = t0 * kS0 + lane;
    float xv = 0.f;
    if (xi < T && xi < len) xv = (row[xi] - mr.x) * mr.y;
    // channel pairs (2 lane + 64 i, + 1) as packed fp32: the 10 x 16 multiply-adds of a frame, the statistics and the
    // normalisation are one f32x2 instruction per pair (this kernel is bound by issue slots: profiles/r02_gelu.md)
    float2 acc[kTT][8];
#pragma unroll
    for (int tt = 0; tt < kTT; ++tt)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[tt][i] = *reinterpret_cast<const float2*>(&b_s[2 * lane + 64 * i]);
#pragma unroll
    for (int j = 0; j < kK0; ++j) {
      float2 wv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wv[i] = *reinterpret_cast<const float2*>(&w_s[j][2 * lane + 64 * i]);
#pragma unroll
      for (int tt = 0; tt < kTT; ++tt) {
        const float2 xs = f2_splat(__shfl_sync(0xffffffffu, xv, tt * kS0 + j));
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[tt][i] = f2_fma(wv[i], xs, acc[tt][i]);
      }
    }
    if (kLayerNorm) {
      // statistics of the kTT frames side by side: the 2 x kTT warp reductions are independent dependency chains (five
      // shuffle + add steps each), kept apart by a per-frame early exit they ran one after the other
      float mean[kTT], rstd[kTT];
#pragma unroll
      for (int tt = 0; tt < kTT; ++tt) {
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s2 = f2_add(s2, acc[tt][i]);
        mean[tt] = s2.x + s2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int tt = 0; tt < kTT; ++tt) mean[tt] += __shfl_xor_sync(0xffffffffu, mean[tt], o);
      }
#pragma unroll
      for (int tt = 0; tt < kTT; ++tt) {
        const float2 neg_mean = f2_splat(-mean[tt] * (1.0f / kC0));
        float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[tt][i] = f2_add(acc[tt][i], neg_mean);
          q2 = f2_fma(acc[tt][i], acc[tt][i], q2);
        }
        rstd[tt] = q2.x + q2.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int tt = 0; tt < kTT; ++tt) rstd[tt] += __shfl_xor_sync(0xffffffffu, rstd[tt], o);
      }
#pragma unroll
      for (int tt = 0; tt < kTT; ++tt) {
        const int t = t0 + tt;
        if (t >= L0) break;  // warp-uniform
        const float2 r2 = f2_splat(rsqrtf(rstd[tt] * (1.0f / kC0) + eps));
        __nv_bfloat16* dst = out + (static_cast<long long>(b) * L0 + t) * kC0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = 2 * lane + 64 * i;
          const float2 g2 = *reinterpret_cast<const float2*>(&g_s[c]);
          const float2 be2 = *reinterpret_cast<const float2*>(&be_s[c]);
          const float2 v = gelu_erf2(f2_fma(acc[tt][i], f2_mul(g2, r2), be2));
          *reinterpret_cast<uint32_t*>(dst + c) = pack_bf16x2(v.x, v.y);
        }
      }
    }
#pragma unroll
    for (int tt = 0; tt < kTT; ++tt) {
      const int t = t0 + tt;
      if (t >= L0) break;  // warp-uniform
      if (kLayerNorm) {
        break;
      } else {
        // GroupNorm(512 groups): statistics run over the whole (padded) time axis per channel
        float* dst = raw_out + (static_cast<long long>(b) * L0 + t) * kC0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = 2 * lane + 64 * i;
          *reinterpret_cast<float2*>(dst + c) = acc[tt][i];
          gsum[2 * i] += acc[tt][i].x;
          gsum[2 * i + 1] += acc[tt][i].y;
          gsq[2 * i] += static_cast<double>(acc[tt][i].x) * acc[tt][i].x;
          gsq[2 * i + 1] += static_cast<double>(acc[tt][i].y) * acc[tt][i].y;
        }
      }
    }
  }
  if (!kLayerNorm) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = 2 * lane + 64 * (i >> 1) + (i & 1);
      atomicAdd(&gn_stats[(static_cast<long long>(b) * kC0 + c) * 2 + 0], gsum[i]);
      atomicAdd(&gn_stats[(static_cast<long long>(b) * kC0 + c) * 2 + 1], gsq[i]);
    }
  }
}

// GroupNorm(num_groups = channels) second pass: normalise each channel with its
// statistics over time, affine, GELU; fp32 -> bf16 channels-last.
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ raw, const double* __restrict__ gn_stats,
                                                              int L0, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              __nv_bfloat16* __restrict__ out) {
  __shared__ float sc_s[kC0], sh_s[kC0];
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < kC0; c += blockDim.x) {
    const double mean = gn_stats[(static_cast<long long>(b) * kC0 + c) * 2 + 0] / L0;
    double var = gn_stats[(static_cast<long long>(b) * kC0 + c) * 2 + 1] / L0 - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float g = gamma ? gamma[c] : 1.f;
    sc_s[c] = rstd * g;
    sh_s[c] = (beta ? beta[c] : 0.f) - static_cast<float>(mean) * rstd * g;
  }
  __syncthreads();
  const long long total = static_cast<long long>(L0) * kC0 / 4;
  const float4* src = reinterpret_cast<const float4*>(raw + static_cast<long long>(b) * L0 * kC0);
  uint2* dst = reinterpret_cast<uint2*>(out + static_cast<long long>(b) * L0 * kC0);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i * 4) % kC0);
    const float4 v = src[i];
    uint2 o;
    o.x = pack_bf16x2(gelu_erf(fmaf(v.x, sc_s[c], sh_s[c])), gelu_erf(fmaf(v.y, sc_s[c + 1], sh_s[c + 1])));
    o.y = pack_bf16x2(gelu_erf(fmaf(v.z, sc_s[c + 2], sh_s[c + 2])), gelu_erf(fmaf(v.w, sc_s[c + 3], sh_s[c + 3])));
    dst[i] = o;
  }
}

// ---------------------------------------------------------------------------
// LayerNorm over the last axis of a row-major matrix, one warp per row.
//   TIn = bf16 or fp32, output bf16 (and optionally fp32), optional GELU.
//   kCols in {512, 1024}.  Two-pass statistics on registers (fp32).
// ---------------------------------------------------------------------------
template <typename TIn, int kCols>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const TIn* __restrict__ in, long long ld_in, long long rows,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float eps, int gelu, __nv_bfloat16* __restrict__ out_bf16,
                                                             long long ld_bf16, float* __restrict__ out_f32,
                                                             long long ld_f32) {
  pdl_trigger();
  pdl_wait();
  constexpr int kPer = kCols / 32;  // elements per lane
  constexpr int kVec = 8;           // elements per vector chunk
  constexpr int kChunks = kPer / kVec;
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  // The row lives in registers as fp32 PAIRS and every per-element step is one packed instruction per pair (add / fma / mul
  // .f32x2): with the GELU these rows are bound by issue slots, not by HBM (profiles/r02_gelu.md), and statistics +
  // normalisation were 6 scalar instructions per element against 2.5 this way.
  float2 v[kPer / 2];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int col = c * 256 + lane * kVec;
    if constexpr (sizeof(TIn) == 2) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + row * ld_in + col);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[c * 4 + i] = __bfloat1622float2(h[i]);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(in + row * ld_in + col);
      const float4 b2 = *reinterpret_cast<const float4*>(in + row * ld_in + col + 4);
      v[c * 4 + 0] = make_float2(a.x, a.y);
      v[c * 4 + 1] = make_float2(a.z, a.w);
      v[c * 4 + 2] = make_float2(b2.x, b2.y);
      v[c * 4 + 3] = make_float2(b2.z, b2.w);
    }
  }
  float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kPer / 2; ++i) s2 = f2_add(s2, v[i]);
  const float mean = warp_sum(s2.x + s2.y) * (1.0f / kCols);
  const float2 neg_mean = f2_splat(-mean);
  float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kPer / 2; ++i) {
    v[i] = f2_add(v[i], neg_mean);  // centred from here on
    q2 = f2_fma(v[i], v[i], q2);
  }
  const float2 rstd = f2_splat(rsqrtf(warp_sum(q2.x + q2.y) * (1.0f / kCols) + eps));
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int col = c * 256 + lane * kVec;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + col);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + col);
    const float4 b1 = *reinterpret_cast<const float4*>(beta + col + 4);
    const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
    const float2 be[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
    float2 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = f2_fma(v[c * 4 + i], f2_mul(g[i], rstd), be[i]);
    if (gelu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = gelu_erf2(o[i]);
    }
    if (out_bf16) {
      uint4 u;
      u.x = pack_bf16x2(o[0].x, o[0].y);
      u.y = pack_bf16x2(o[1].x, o[1].y);
      u.z = pack_bf16x2(o[2].x, o[2].y);
      u.w = pack_bf16x2(o[3].x, o[3].y);
      *reinterpret_cast<uint4*>(out_bf16 + row * ld_bf16 + col) = u;
    }
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + row * ld_f32 + col) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
      *reinterpret_cast<float4*>(out_f32 + row * ld_f32 + col + 4) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
    }
  }
}

static inline int grid_for(long long work_items, int per_block, int max_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return static_cast<int>(g);
}

}  // namespace aph

using namespace aph;

extern "C" int aph_wave_stats(const float* x, const int64_t* lengths, int32_t n_utt, int32_t T, double* stats_scratch,
                              float* mean_rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && lengths && stats_scratch && mean_rstd, "null pointer");
  APH_REQUIRE(n_utt > 0 && T > 0, "empty batch");
  APH_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 3 * n_utt, stream));
  const int chunks = grid_for(T, 256 * 16, ceil_div(4 * sm_count(), n_utt) > 0 ? ceil_div(4 * sm_count(), n_utt) : 1);
  APH_CUDA_CHECK(launch_pdl(wave_stats_kernel, dim3(chunks, n_utt), dim3(256), 0, stream, x, reinterpret_cast<const long long*>(lengths), T,
                            stats_scratch));
  APH_CUDA_CHECK(launch_pdl(wave_finalize_kernel, dim3(ceil_div(n_utt, 128)), dim3(128), 0, stream, stats_scratch,
                            reinterpret_cast<const long long*>(lengths), n_utt, reinterpret_cast<float2*>(mean_rstd)));
  APH_POST_LAUNCH(2);
  return APH_OK;
}

extern "C" int aph_wave_norm(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt, int32_t T,
                             float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && lengths && mean_rstd && out, "null pointer");
  APH_REQUIRE(n_utt > 0 && T > 0, "empty batch");
  const int gx = grid_for(T, 256 * 8, 1024);
  wave_norm_kernel<<<dim3(gx, n_utt), 256, 0, stream>>>(x, reinterpret_cast<const long long*>(lengths),
                                                       reinterpret_cast<const float2*>(mean_rstd), T, out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_frame_lengths(const int64_t* lengths, int32_t n_utt, const int32_t* kernels, const int32_t* strides,
                                 int32_t n_layers, int32_t* frames32, int64_t* frames64, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(lengths && kernels && strides, "null pointer");
  APH_REQUIRE(n_utt > 0 && n_layers >= 0, "empty batch");
  APH_CUDA_CHECK(launch_pdl(frame_lengths_kernel, dim3(ceil_div(n_utt, 128)), dim3(128), 0, stream,
                            reinterpret_cast<const long long*>(lengths), n_utt, kernels, strides, n_layers, frames32,
                            reinterpret_cast<long long*>(frames64)));
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_conv0_ln_gelu(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt, int32_t T,
                                 const float* w, const float* bias, const float* gamma, const float* beta, float eps,
                                 int32_t skip_padded_frames, void* out_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && w && gamma && beta && out_bf16, "null pointer");
  APH_REQUIRE(n_utt > 0 && T >= kK0, "waveform shorter than the first conv kernel");
  const int L0 = (T - kK0) / kS0 + 1;
  const int gx = grid_for(L0, 8 * kTT * 4, ceil_div(8 * sm_count(), n_utt) > 0 ? ceil_div(8 * sm_count(), n_utt) : 1);
  APH_CUDA_CHECK(launch_pdl(conv0_kernel<true>, dim3(gx, n_utt), dim3(256), 0, stream, x, reinterpret_cast<const long long*>(lengths),
                            reinterpret_cast<const float2*>(mean_rstd), T, L0, w, bias, gamma, beta, eps,
                            static_cast<__nv_bfloat16*>(out_bf16), static_cast<float*>(nullptr), static_cast<double*>(nullptr),
                            skip_padded_frames));
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_conv0_gn_gelu(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt, int32_t T,
                                 const float* w, const float* bias, const float* gamma, const float* beta, float eps,
                                 float* raw_scratch, double* stats_scratch, void* out_bf16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && w && raw_scratch && stats_scratch && out_bf16, "null pointer");
  APH_REQUIRE(n_utt > 0 && T >= kK0, "waveform shorter than the first conv kernel");
  const int L0 = (T - kK0) / kS0 + 1;
  APH_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 2 * kC0 * n_utt, stream));
  const int gx = grid_for(L0, 8 * kTT * 16, ceil_div(4 * sm_count(), n_utt) > 0 ? ceil_div(4 * sm_count(), n_utt) : 1);
  APH_CUDA_CHECK(launch_pdl(conv0_kernel<false>, dim3(gx, n_utt), dim3(256), 0, stream, x, reinterpret_cast<const long long*>(lengths),
                            reinterpret_cast<const float2*>(mean_rstd), T, L0, w, bias, static_cast<const float*>(nullptr),
                            static_cast<const float*>(nullptr), eps, static_cast<__nv_bfloat16*>(nullptr), raw_scratch, stats_scratch, 0));
  const int gy = grid_for(static_cast<long long>(L0) * kC0 / 4, 256 * 8, 2048);
  groupnorm_apply_kernel<<<dim3(gy, n_utt), 256, 0, stream>>>(raw_scratch, stats_scratch, L0, gamma, beta, eps,
                                                             static_cast<__nv_bfloat16*>(out_bf16));
  APH_POST_LAUNCH(2);
  return APH_OK;
}

extern "C" int aph_layernorm_rows(const void* in, int32_t in_is_f32, int64_t ld_in, int64_t rows, int32_t cols,
                                  const float* gamma, const float* beta, float eps, int32_t gelu, void* out_bf16,
                                  int64_t ld_bf16, float* out_f32, int64_t ld_f32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(in && gamma && beta && (out_bf16 || out_f32), "null pointer");
  APH_REQUIRE(cols == 512 || cols == 1024, "layernorm supports 512 or 1024 columns");
  APH_REQUIRE(ld_in % 8 == 0 && ld_bf16 % 8 == 0 && ld_f32 % 4 == 0, "leading dimensions must keep 16-byte alignment");
  if (rows <= 0) return APH_OK;
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>((rows + warps - 1) / warps);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
  if (in_is_f32) {
    const float* p = static_cast<const float*>(in);
    if (cols == 512)
      APH_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<float, 512>, dim3(grid), dim3(256), 0, stream, p, ld_in, rows, gamma, beta, eps, gelu, ob, ld_bf16, out_f32, ld_f32));
    else
      APH_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<float, 1024>, dim3(grid), dim3(256), 0, stream, p, ld_in, rows, gamma, beta, eps, gelu, ob, ld_bf16, out_f32, ld_f32));
  } else {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(in);
    if (cols == 512)
      APH_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<__nv_bfloat16, 512>, dim3(grid), dim3(256), 0, stream, p, ld_in, rows, gamma, beta, eps, gelu, ob, ld_bf16, out_f32, ld_f32));
    else
      APH_CUDA_CHECK(launch_pdl(layernorm_rows_kernel<__nv_bfloat16, 1024>, dim3(grid), dim3(256), 0, stream, p, ld_in, rows, gamma, beta, eps, gelu, ob, ld_bf16, out_f32, ld_f32));
  }
  APH_POST_LAUNCH(1);
  return APH_OK;
}
