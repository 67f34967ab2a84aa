// Training of the wav2vec2 convolutional feature extractor (freeze_feature_encoder = false, or after the reference's
// UnfreezeSchedule fired; HF modeling_wav2vec2.py:275-323, 382-419, layer-norm variant): the pieces the frozen-extractor
// path fuses away or never needed.
//   conv0_raw_kernel             conv(1 -> 512, k 10, s 5) of the normalised waveform, PRE-LayerNorm, bf16 (kept for the backward)
//   ln_gelu_backward_512_kernel  backward of (LayerNorm(512) -> GELU) from the kept pre-LayerNorm conv output: recomputes the
//                                normalised value, applies GELU', LayerNorm backward; dgamma / dbeta reduced per block
//   conv0_weight_backward_kernel dW0[o][j] = sum_(n,t) dY[n][t][o] * xn[n][5 t + j], db0[o] = sum dY
// The strided convolutions 1..6 reuse the GEMM (weight gradient over overlapping-row windows, data gradient) and
// aph_conv_input_backward (col2im).  All HBM-bound; this path is the reference's non-default configuration.
#include "aph_common.cuh"

namespace aph {

constexpr int kXC = 512;
constexpr int kXK0 = 10;
constexpr int kXS0 = 5;

__global__ void __launch_bounds__(256) conv0_raw_kernel(const float* __restrict__ x, const long long* __restrict__ lengths,
                                                        const float2* __restrict__ mean_rstd, int T, int L0,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ out) {
  __shared__ float w_s[kXK0][kXC];
  __shared__ float b_s[kXC];
  __shared__ float win[8][16];
  for (int i = threadIdx.x; i < kXC * kXK0; i += blockDim.x) w_s[i % kXK0][i / kXK0] = w[i];
  for (int i = threadIdx.x; i < kXC; i += blockDim.x) b_s[i] = bias ? bias[i] : 0.f;
  const int n = blockIdx.y;
  const long long len = lengths ? lengths[n] : T;
  const float2 mr = mean_rstd ? mean_rstd[n] : make_float2(0.f, 1.f);
  const float* row = x + static_cast<long long>(n) * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  for (int t = blockIdx.x * 8 + warp; t < L0; t += gridDim.x * 8) {
    if (lane < kXK0) {
      const int xi = t * kXS0 + lane;
      win[warp][lane] = (xi < T && xi < len) ? (row[xi] - mr.x) * mr.y : 0.f;
    }
    __syncwarp();
    __nv_bfloat16* dst = out + (static_cast<long long>(n) * L0 + t) * kXC;
#pragma unroll 4
    for (int c = lane; c < kXC; c += 32) {
      float acc = b_s[c];
#pragma unroll
      for (int j = 0; j < kXK0; ++j) acc = fmaf(w_s[j][c], win[warp][j], acc);
      dst[c] = __float2bfloat16(acc);
    }
    __syncwarp();
  }
}

// x: pre-LayerNorm conv output bf16 [rows][512]; d_out: gradient of the GELU output fp32 [rows][ld_d];
// dx: gradient of the conv output, bf16 [rows][512] (operand of the conv's gradient GEMMs); dgamma / dbeta / dbias accumulate.
__global__ void __launch_bounds__(256) ln_gelu_backward_512_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ d_out,
                                                                   long long ld_d, long long rows, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, float eps,
                                                                   __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                   float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ float red[3][8][kXC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float pg[16], pb[16], px[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) pg[i] = pb[i] = px[i] = 0.f;
  float gm[16], bt[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    gm[i] = gamma[lane + 32 * i];
    bt[i] = beta[lane + 32 * i];
  }
  for (long long row = blockIdx.x * 8ll + warp; row < rows; row += gridDim.x * 8ll) {
    float v[16], d[16];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] = __bfloat162float(x[row * kXC + lane + 32 * i]);
      d[i] = d_out[row * ld_d + lane + 32 * i];
      sum += v[i];
    }
    const float mean = warp_sum(sum) * (1.0f / kXC);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] -= mean;
      sq = fmaf(v[i], v[i], sq);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / kXC) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      v[i] *= rstd;                                          // xhat
      const float dy = d[i] * gelu_erf_grad(fmaf(v[i], gm[i], bt[i]));  // through the GELU
      pg[i] = fmaf(dy, v[i], pg[i]);
      pb[i] += dy;
      d[i] = dy * gm[i];
      s1 += d[i];
      s2 = fmaf(d[i], v[i], s2);
    }
    const float c1 = warp_sum(s1) * (1.0f / kXC), c2 = warp_sum(s2) * (1.0f / kXC);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float g = rstd * (d[i] - c1 - v[i] * c2);
      px[i] += g;
      dx[row * kXC + lane + 32 * i] = __float2bfloat16(g);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    red[0][warp][lane + 32 * i] = pg[i];
    red[1][warp][lane + 32 * i] = pb[i];
    red[2][warp][lane + 32 * i] = px[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kXC; c += blockDim.x) {
    float a = 0.f, b = 0.f, e = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a += red[0][k][c];
      b += red[1][k][c];
      e += red[2][k][c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
    if (dbias != nullptr) atomicAdd(dbias + c, e);
  }
}

// dW0[o][j] += sum over this block's frames of dY[n][t][o] * xn[n][5 t + j]   (dY bf16 [N][L0][512])
__global__ void __launch_bounds__(512) conv0_weight_backward_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                                    const long long* __restrict__ lengths,
                                                                    const float2* __restrict__ mean_rstd, int T, int L0,
                                                                    float* __restrict__ dw /*[512][10]*/) {
  constexpr int kFrames = 64;
  __shared__ float xs[kFrames * kXS0 + kXK0];
  const int n = blockIdx.y;
  const int o = threadIdx.x;
  const long long len = lengths ? lengths[n] : T;
  const float2 mr = mean_rstd ? mean_rstd[n] : make_float2(0.f, 1.f);
  const float* row = x + static_cast<long long>(n) * T;
  float acc[kXK0];
#pragma unroll
  for (int j = 0; j < kXK0; ++j) acc[j] = 0.f;
  for (int t0 = blockIdx.x * kFrames; t0 < L0; t0 += gridDim.x * kFrames) {
    __syncthreads();
    for (int i = threadIdx.x; i < kFrames * kXS0 + kXK0; i += blockDim.x) {
      const int xi = t0 * kXS0 + i;
      xs[i] = (xi < T && xi < len) ? (row[xi] - mr.x) * mr.y : 0.f;
    }
    __syncthreads();
    const int nt = min(kFrames, L0 - t0);
    for (int tt = 0; tt < nt; ++tt) {
      const float g = __bfloat162float(dy[(static_cast<long long>(n) * L0 + t0 + tt) * kXC + o]);
#pragma unroll
      for (int j = 0; j < kXK0; ++j) acc[j] = fmaf(g, xs[tt * kXS0 + j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kXK0; ++j) atomicAdd(dw + o * kXK0 + j, acc[j]);
}

}  // namespace aph

using namespace aph;

extern "C" int aph_conv0_raw_bf16(const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt, int32_t T, const float* w,
                                  const float* bias, void* out_bf16, void* stream_) {
  APH_REQUIRE(x && w && out_bf16 && n_utt > 0 && T >= kXK0, "conv0_raw: bad arguments");
  APH_REQUIRE(n_utt <= 65535, "conv0_raw: at most 65535 utterances");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L0 = (T - kXK0) / kXS0 + 1;
  const int gx = std::max(1, std::min((L0 + 7) / 8, 4 * sm_count() / n_utt + 1));
  conv0_raw_kernel<<<dim3(gx, n_utt), 256, 0, stream>>>(x, reinterpret_cast<const long long*>(lengths), reinterpret_cast<const float2*>(mean_rstd), T, L0,
                                                        w, bias, static_cast<__nv_bfloat16*>(out_bf16));
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_ln_gelu_backward_512(const void* x_bf16, const float* d_out, int64_t ld_d, int64_t rows, const float* gamma,
                                        const float* beta, float eps, void* dx_bf16, float* dgamma, float* dbeta, float* dbias,
                                        void* stream_) {
  APH_REQUIRE(x_bf16 && d_out && gamma && beta && dx_bf16 && dgamma && dbeta && rows >= 0 && ld_d >= kXC, "ln_gelu_backward_512: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_CUDA_CHECK(cudaMemsetAsync(dgamma, 0, sizeof(float) * kXC, stream));
  APH_CUDA_CHECK(cudaMemsetAsync(dbeta, 0, sizeof(float) * kXC, stream));
  if (dbias) APH_CUDA_CHECK(cudaMemsetAsync(dbias, 0, sizeof(float) * kXC, stream));
  if (rows == 0) return APH_OK;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((rows + 7) / 8, 4ll * sm_count()));
  ln_gelu_backward_512_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x_bf16), d_out, ld_d, rows, gamma, beta, eps,
                                                       static_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dbias);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_conv0_weight_backward(const void* dy_bf16, const float* x, const int64_t* lengths, const float* mean_rstd, int32_t n_utt,
                                         int32_t T, float* dw, void* stream_) {
  APH_REQUIRE(dy_bf16 && x && dw && n_utt > 0 && T >= kXK0, "conv0_weight_backward: bad arguments");
  APH_REQUIRE(n_utt <= 65535, "conv0_weight_backward: at most 65535 utterances");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int L0 = (T - kXK0) / kXS0 + 1;
  APH_CUDA_CHECK(cudaMemsetAsync(dw, 0, sizeof(float) * kXC * kXK0, stream));
  const int gx = std::max(1, std::min((L0 + 63) / 64, 4 * sm_count() / n_utt + 1));
  conv0_weight_backward_kernel<<<dim3(gx, n_utt), 512, 0, stream>>>(static_cast<const __nv_bfloat16*>(dy_bf16), x,
                                                                    reinterpret_cast<const long long*>(lengths),
                                                                    reinterpret_cast<const float2*>(mean_rstd), T, L0, dw);
  APH_POST_LAUNCH(1);
  return APH_OK;
}
