// allophant_b200 — optimiser step of the training loop (estimator.py:778-791): global-norm gradient clipping
// (nn.utils.clip_grad_norm_), torch.optim.Adam (betas 0.9 / 0.98 in the reference's default config, config.py:316-335)
// as multi-tensor kernels: the per-tensor descriptors travel by value in the kernel parameters, one launch covers up
// to kMtTensors tensors (the reference issues a few small kernels per parameter tensor: ~500 tensors x ~10 launches).
//
//   aph_multi_tensor_sumsq   sum over all tensors of x^2, accumulated in fp64 -> *out (double)
//   aph_multi_tensor_scale   x *= min(1, max_norm / (sqrt(*sumsq) + 1e-6))        (clip_grad_norm_, in place)
//   aph_multi_tensor_adam    p, exp_avg, exp_avg_sq <- Adam update; optional bf16 shadow copy of p (the GEMM operand)
//
// Memory-bound: Adam reads 4 fp32 streams and writes 3 (+ 2 bytes per element of shadow weights).
#include <string.h>

#include "aph_common.cuh"

namespace aph {

constexpr int kMtTensors = 48;
constexpr int kMtBlock = 512;
constexpr int kMtChunk = 8192;  // elements per block iteration unit

struct MtPack {
  void* p[kMtTensors][4];  // up to four pointers per tensor (role depends on the kernel)
  long long n[kMtTensors];
  int chunk_start[kMtTensors + 1];  // prefix sum of ceil(n / kMtChunk): block b works on the tensor owning chunk b
  int count;
};

__device__ __forceinline__ int mt_find(const MtPack& pack, int chunk) {
  int lo = 0, hi = pack.count;  // largest t with chunk_start[t] <= chunk
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pack.chunk_start[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kMtBlock) mt_sumsq_kernel(const __grid_constant__ MtPack pack, double* __restrict__ out) {
  __shared__ double red[kMtBlock / 32];
  double acc = 0.0;
  const int total_chunks = pack.chunk_start[pack.count];
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int t = mt_find(pack, chunk);
    const float* x = static_cast<const float*>(pack.p[t][0]);
    const long long base = static_cast<long long>(chunk - pack.chunk_start[t]) * kMtChunk;
    const long long end = min(pack.n[t], base + kMtChunk);
    float part = 0.f;
    for (long long i = base + threadIdx.x; i < end; i += kMtBlock) {
      const float v = x[i];
      part = fmaf(v, v, part);
    }
    acc += static_cast<double>(part);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kMtBlock / 32 ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

__device__ __forceinline__ float clip_coefficient(const double* sumsq, float max_norm) {
  if (sumsq == nullptr || !(max_norm > 0.f)) return 1.f;
  const float norm = static_cast<float>(sqrt(*sumsq));
  const float coef = max_norm / (norm + 1e-6f);  // nn.utils.clip_grad_norm_: clamped to at most 1
  return coef < 1.f ? coef : 1.f;
}

__global__ void __launch_bounds__(kMtBlock) mt_scale_kernel(const __grid_constant__ MtPack pack, const double* __restrict__ sumsq,
                                                            float max_norm) {
  const float coef = clip_coefficient(sumsq, max_norm);
  if (coef == 1.f) return;
  const int total_chunks = pack.chunk_start[pack.count];
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int t = mt_find(pack, chunk);
    float* x = static_cast<float*>(pack.p[t][0]);
    const long long base = static_cast<long long>(chunk - pack.chunk_start[t]) * kMtChunk;
    const long long end = min(pack.n[t], base + kMtChunk);
    for (long long i = base + threadIdx.x; i < end; i += kMtBlock) x[i] *= coef;
  }
}

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  float bias_correction1;       // 1 - beta1^t
  float bias_correction2_sqrt;  // sqrt(1 - beta2^t)
  float max_norm;               // > 0: gradients are scaled by the clip coefficient on the fly (grad buffers untouched)
};

// torch.optim.Adam (no amsgrad, not maximize): g += wd * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(kMtBlock) mt_adam_kernel(const __grid_constant__ MtPack pack, const AdamHyper h,
                                                           const double* __restrict__ sumsq,
                                                           const __grid_constant__ MtPack shadow) {
  const float coef = clip_coefficient(sumsq, h.max_norm);
  const float step_size = h.lr / h.bias_correction1;
  const int total_chunks = pack.chunk_start[pack.count];
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int t = mt_find(pack, chunk);
    float* p = static_cast<float*>(pack.p[t][0]);
    const float* g = static_cast<const float*>(pack.p[t][1]);
    float* m = static_cast<float*>(pack.p[t][2]);
    float* v = static_cast<float*>(pack.p[t][3]);
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(shadow.p[t][0]);
    const long long base = static_cast<long long>(chunk - pack.chunk_start[t]) * kMtChunk;
    const long long end = min(pack.n[t], base + kMtChunk);
    auto update = [&](float pi, float gi, float& mi, float& vi) -> float {
      gi = fmaf(h.weight_decay, pi, gi * coef);
      mi = fmaf(h.beta1, mi, (1.f - h.beta1) * gi);
      vi = fmaf(h.beta2, vi, (1.f - h.beta2) * gi * gi);
      const float denom = sqrtf(vi) / h.bias_correction2_sqrt + h.eps;
      return pi - step_size * (mi / denom);
    };
    // chunks start at multiples of 8192 elements: 16-byte accesses whenever the tensors themselves are 16-byte aligned
    const bool wide = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sh) & 7) == 0;
    long long done = base;
    if (wide) {
      const long long quads = (end - base) >> 2;
      for (long long q = threadIdx.x; q < quads; q += kMtBlock) {
        const long long i = base + 4 * q;
        float4 pq = *reinterpret_cast<const float4*>(p + i);
        const float4 gq = __ldcs(reinterpret_cast<const float4*>(g + i));  // the gradient is not read again
        float4 mq = *reinterpret_cast<const float4*>(m + i);
        float4 vq = *reinterpret_cast<const float4*>(v + i);
        pq.x = update(pq.x, gq.x, mq.x, vq.x);
        pq.y = update(pq.y, gq.y, mq.y, vq.y);
        pq.z = update(pq.z, gq.z, mq.z, vq.z);
        pq.w = update(pq.w, gq.w, mq.w, vq.w);
        *reinterpret_cast<float4*>(m + i) = mq;
        *reinterpret_cast<float4*>(v + i) = vq;
        *reinterpret_cast<float4*>(p + i) = pq;
        if (sh != nullptr) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(pq.x, pq.y), hi = __floats2bfloat162_rn(pq.z, pq.w);
          uint2 packed;
          packed.x = *reinterpret_cast<const uint32_t*>(&lo);
          packed.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(sh + i) = packed;
        }
      }
      done = base + 4 * quads;
    }
    for (long long i = done + threadIdx.x; i < end; i += kMtBlock) {
      float mi = m[i], vi = v[i];
      const float pn = update(p[i], g[i], mi, vi);
      m[i] = mi;
      v[i] = vi;
      p[i] = pn;
      if (sh != nullptr) sh[i] = __float2bfloat16(pn);
    }
  }
}

// torch.optim.SGD (dampening 0, no Nesterov): g += wd * p; buf = first ? g : momentum * buf + g; p -= lr * (momentum ? buf : g)
__global__ void __launch_bounds__(kMtBlock) mt_sgd_kernel(const __grid_constant__ MtPack pack, float lr, float momentum, float weight_decay,
                                                          int first_step, float max_norm, const double* __restrict__ sumsq,
                                                          const __grid_constant__ MtPack shadow) {
  const float coef = clip_coefficient(sumsq, max_norm);
  const int total_chunks = pack.chunk_start[pack.count];
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int t = mt_find(pack, chunk);
    float* p = static_cast<float*>(pack.p[t][0]);
    const float* g = static_cast<const float*>(pack.p[t][1]);
    float* buf = static_cast<float*>(pack.p[t][2]);
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(shadow.p[t][0]);
    const long long base = static_cast<long long>(chunk - pack.chunk_start[t]) * kMtChunk;
    const long long end = min(pack.n[t], base + kMtChunk);
    for (long long i = base + threadIdx.x; i < end; i += kMtBlock) {
      const float pi = p[i];
      float gi = fmaf(weight_decay, pi, g[i] * coef);
      if (buf != nullptr) {
        gi = first_step ? gi : fmaf(momentum, buf[i], gi);
        buf[i] = gi;
      }
      const float pn = fmaf(-lr, gi, pi);
      p[i] = pn;
      if (sh != nullptr) sh[i] = __float2bfloat16(pn);
    }
  }
}

static int build_pack(MtPack& pack, void* const* ptrs, int roles, const int64_t* sizes, int first, int count) {
  memset(&pack, 0, sizeof(pack));
  pack.count = count;
  int chunks = 0;
  for (int t = 0; t < count; ++t) {
    for (int r = 0; r < roles; ++r) pack.p[t][r] = ptrs ? ptrs[static_cast<size_t>(first + t) * roles + r] : nullptr;
    pack.n[t] = sizes[first + t];
    pack.chunk_start[t] = chunks;
    chunks += static_cast<int>((sizes[first + t] + kMtChunk - 1) / kMtChunk);
  }
  pack.chunk_start[count] = chunks;
  return chunks;
}

static unsigned mt_grid(int chunks) {
  const int cap = 8 * sm_count();
  const int g = chunks < cap ? chunks : cap;
  return static_cast<unsigned>(g < 1 ? 1 : g);
}

}  // namespace aph

using namespace aph;

extern "C" int aph_multi_tensor_sumsq(void* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors, double* out,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(tensors_host && sizes_host && out && n_tensors >= 0, "null pointer");
  APH_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double), stream));
  int launched = 0;
  for (int first = 0; first < n_tensors; first += kMtTensors) {
    const int count = n_tensors - first < kMtTensors ? n_tensors - first : kMtTensors;
    MtPack pack;
    const int chunks = build_pack(pack, tensors_host, 1, sizes_host, first, count);
    if (chunks == 0) continue;
    mt_sumsq_kernel<<<mt_grid(chunks), kMtBlock, 0, stream>>>(pack, out);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_multi_tensor_scale(void* const* tensors_host, const int64_t* sizes_host, int32_t n_tensors,
                                      const double* sumsq, float max_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(tensors_host && sizes_host && sumsq && n_tensors >= 0, "null pointer");
  int launched = 0;
  for (int first = 0; first < n_tensors; first += kMtTensors) {
    const int count = n_tensors - first < kMtTensors ? n_tensors - first : kMtTensors;
    MtPack pack;
    const int chunks = build_pack(pack, tensors_host, 1, sizes_host, first, count);
    if (chunks == 0) continue;
    mt_scale_kernel<<<mt_grid(chunks), kMtBlock, 0, stream>>>(pack, sumsq, max_norm);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_multi_tensor_adam(void* const* tensors_host /*[n][4]: param, grad, exp_avg, exp_avg_sq*/,
                                     void* const* shadow_bf16_host /*[n] or NULL*/, const int64_t* sizes_host, int32_t n_tensors,
                                     float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                     const double* clip_sumsq, float max_norm, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(tensors_host && sizes_host && n_tensors >= 0 && step >= 1, "bad arguments");
  AdamHyper h;
  h.lr = lr;
  h.beta1 = beta1;
  h.beta2 = beta2;
  h.eps = eps;
  h.weight_decay = weight_decay;
  h.bias_correction1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  h.bias_correction2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  h.max_norm = clip_sumsq ? max_norm : 0.f;
  int launched = 0;
  for (int first = 0; first < n_tensors; first += kMtTensors) {
    const int count = n_tensors - first < kMtTensors ? n_tensors - first : kMtTensors;
    MtPack pack, shadow;
    const int chunks = build_pack(pack, tensors_host, 4, sizes_host, first, count);
    build_pack(shadow, shadow_bf16_host, 1, sizes_host, first, count);
    if (chunks == 0) continue;
    mt_adam_kernel<<<mt_grid(chunks), kMtBlock, 0, stream>>>(pack, h, clip_sumsq, shadow);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_multi_tensor_sgd(void* const* tensors_host /*[n][3]: param, grad, momentum buffer (or NULL)*/,
                                    void* const* shadow_bf16_host /*[n] or NULL*/, const int64_t* sizes_host, int32_t n_tensors, float lr,
                                    float momentum, float weight_decay, int32_t first_step, const double* clip_sumsq, float max_norm,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(tensors_host && sizes_host && n_tensors >= 0, "bad arguments");
  int launched = 0;
  for (int first = 0; first < n_tensors; first += kMtTensors) {
    const int count = n_tensors - first < kMtTensors ? n_tensors - first : kMtTensors;
    MtPack pack, shadow;
    const int chunks = build_pack(pack, tensors_host, 3, sizes_host, first, count);
    build_pack(shadow, shadow_bf16_host, 1, sizes_host, first, count);
    if (chunks == 0) continue;
    mt_sgd_kernel<<<mt_grid(chunks), kMtBlock, 0, stream>>>(pack, lr, momentum, weight_decay, first_step, clip_sumsq ? max_norm : 0.f, clip_sumsq, shadow);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}
