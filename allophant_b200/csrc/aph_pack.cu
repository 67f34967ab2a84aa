// allophant_b200 — weight packing kernels (run once per weight version, not per step
// in inference): fp32 checkpoint tensors -> bf16 operands in the layouts the GEMM wants.
//
//   aph_cast_bf16            nn.Linear weights ([out][in] is already the K-major B operand)
//   aph_pack_conv_weight     Conv1d weight [O][C][k] -> [O][k][C] so that k-block j of the
//                            implicit GEMM is the contiguous window of channels-last input
//   aph_pack_posconv_weight  weight_norm(dim=2) of the positional conv (HF:326-350):
//                            w[o][c][j] = g[j] * v[o][c][j] / ||v[:, :, j]||_2, then -> [O][k][Cg]
#include "aph_common.cuh"

namespace aph {

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16(src[i]);
}

// strided 2-D cast: dst[r][c] = bf16(src[r][c]) with independent leading dimensions
__global__ void __launch_bounds__(256) cast_bf16_2d_kernel(const float* __restrict__ src, long long ld_src,
                                                           __nv_bfloat16* __restrict__ dst, long long ld_dst, long long rows,
                                                           int cols4) {
  const long long total = rows * cols4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = o;
  }
}

__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                               int O, int C, int k, const float* __restrict__ tap_scale) {
  const long long total = static_cast<long long>(O) * C * k;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    // i indexes dst [o][j][c]
    const int c = static_cast<int>(i % C);
    const int j = static_cast<int>((i / C) % k);
    const long long o = i / (static_cast<long long>(C) * k);
    float v = src[(o * C + c) * k + j];
    if (tap_scale) v *= tap_scale[j];
    dst[i] = __float2bfloat16(v);
  }
}

// tap_scale[j] = g[j] / sqrt(sum_{o,c} v[o][c][j]^2)
__global__ void __launch_bounds__(256) posconv_tap_scale_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                                long long oc, int k, float* __restrict__ tap_scale) {
  const int j = blockIdx.x;
  double s = 0.0;
  for (long long i = threadIdx.x; i < oc; i += blockDim.x) {
    const float x = v[i * k + j];
    s += static_cast<double>(x) * x;
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) tap_scale[j] = static_cast<float>(static_cast<double>(g[j]) / sqrt(red[0]));
}

// LayerNorm folded into the Linear behind it: Linear(LayerNorm(x)) = rstd * (x W'^T - mean * colsum) + bias' with
//   W'[n][k] = W[n][k] * gamma[k] (bf16 GEMM operand),  colsum[n] = sum_k W'[n][k] (of the ROUNDED operand, so that a row
//   of equal values cancels exactly),  bias'[n] = bias[n] + sum_k W[n][k] * beta[k].
// One block per output feature.
__global__ void __launch_bounds__(256) fold_layernorm_linear_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    int k, __nv_bfloat16* __restrict__ out_w, float* __restrict__ out_colsum,
                                                                    float* __restrict__ out_bias) {
  const int n = blockIdx.x;
  const float* row = w + static_cast<long long>(n) * k;
  __nv_bfloat16* dst = out_w + static_cast<long long>(n) * k;
  float cs = 0.f, bs = 0.f;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const float wv = row[i];
    const __nv_bfloat16 folded = __float2bfloat16(wv * gamma[i]);
    dst[i] = folded;
    cs += __bfloat162float(folded);
    bs = fmaf(wv, beta[i], bs);
  }
  __shared__ float red[2][8];
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = cs;
    red[1][warp] = bs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      b += red[1][i];
    }
    out_colsum[n] = a;
    out_bias[n] = (bias ? bias[n] : 0.f) + b;
  }
}

}  // namespace aph

using namespace aph;

extern "C" int aph_fold_layernorm_linear(const float* weight, const float* bias, const float* gamma, const float* beta, int32_t n,
                                         int32_t k, void* out_weight_bf16, float* out_colsum, float* out_bias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(weight && gamma && beta && out_weight_bf16 && out_colsum && out_bias, "null pointer");
  APH_REQUIRE(n > 0 && k > 0, "bad shape");
  fold_layernorm_linear_kernel<<<n, 256, 0, stream>>>(weight, bias, gamma, beta, k, static_cast<__nv_bfloat16*>(out_weight_bf16),
                                                      out_colsum, out_bias);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(src && dst_bf16, "null pointer");
  if (n <= 0) return APH_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  cast_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst_bf16), n);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_cast_bf16_2d(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int64_t rows, int32_t cols,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(src && dst_bf16, "null pointer");
  APH_REQUIRE(cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0, "columns and leading dimensions must be multiples of 4");
  if (rows <= 0 || cols <= 0) return APH_OK;
  long long blocks = (rows * (cols / 4) + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  cast_bf16_2d_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, ld_src, static_cast<__nv_bfloat16*>(dst_bf16), ld_dst,
                                                                         rows, cols / 4);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_pack_conv_weight(const float* src, void* dst_bf16, int32_t out_channels, int32_t in_channels,
                                    int32_t kernel, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(src && dst_bf16, "null pointer");
  APH_REQUIRE(out_channels > 0 && in_channels > 0 && kernel > 0, "bad shape");
  long long blocks = (static_cast<long long>(out_channels) * in_channels * kernel + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  pack_conv_weight_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst_bf16),
                                                                             out_channels, in_channels, kernel, nullptr);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_pack_posconv_weight(const float* weight_g, const float* weight_v, void* dst_bf16, float* tap_scale_scratch,
                                       int32_t out_channels, int32_t group_channels, int32_t kernel, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(weight_g && weight_v && dst_bf16 && tap_scale_scratch, "null pointer");
  APH_REQUIRE(out_channels > 0 && group_channels > 0 && kernel > 0, "bad shape");
  posconv_tap_scale_kernel<<<kernel, 256, 0, stream>>>(weight_v, weight_g, static_cast<long long>(out_channels) * group_channels,
                                                       kernel, tap_scale_scratch);
  long long blocks = (static_cast<long long>(out_channels) * group_channels * kernel + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  pack_conv_weight_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(weight_v, static_cast<__nv_bfloat16*>(dst_bf16),
                                                                             out_channels, group_channels, kernel,
                                                                             tap_scale_scratch);
  APH_POST_LAUNCH(2);
  return APH_OK;
}
