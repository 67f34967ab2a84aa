// Micro-benchmark (one B200): issue cost of the candidate softmax instructions of the attention kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu_bench tools/micro/mufu_bench.cu && /tmp/mufu_bench
// Each kernel runs `kIters` dependent-free instructions per thread in 8 independent chains; one CTA of `warps` warps per SM.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

constexpr int kIters = 4096;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void bench(float* out, long long* cycles) {
  float f[8];
  uint32_t u[8];
  uint64_t w[8];
  for (int i = 0; i < 8; ++i) { f[i] = -0.001f * (threadIdx.x + i); u[i] = 0xBF80BF80u + i; w[i] = 0x3F8000003F800000ull + i; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) f[i] = ex2f(f[i]);
      if (MODE == 1) u[i] = ex2bf2(u[i]);
      if (MODE == 2) u[i] = ex2h2(u[i]);
      if (MODE == 3) f[i] = max3(f[i], f[(i + 1) & 7], -1.0f);
      if (MODE == 4) f[i] = fmaxf(f[i], -1.0f + f[(i + 1) & 7]);
      if (MODE == 5) w[i] = fma2(w[i], w[(i + 1) & 7], w[i]);
      if (MODE == 6) w[i] = add2(w[i], w[(i + 1) & 7]);
      if (MODE == 7) f[i] = fmaf(f[i], 0.999f, 0.001f);
      if (MODE == 8) { float y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(f[i]), "f"(f[(i + 1) & 7])); }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += f[i] + __uint_as_float(u[i]) + __uint_as_float(static_cast<uint32_t>(w[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void fence_bench(float* out, long long* cycles) {
  __shared__ float buf[1024];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < 256; ++it) {
    buf[threadIdx.x] = it;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const long long t1 = clock64();
  out[threadIdx.x] = buf[threadIdx.x];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (t1 - t0);
}

template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, sizeof(float) * 148 * 1024); cudaMalloc(&cyc, 8);
  bench<MODE><<<148, warps * 32>>>(out, cyc);
  bench<MODE><<<148, warps * 32>>>(out, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // cycles per warp-instruction per SMSP: warps/4 warps share one scheduler
  printf("%-34s warps/SM %2d: %7.2f cycles per warp-instruction per SMSP (%lld cycles, %s)\n", name, warps, (double)h / kIters / (warps / 4.0), h, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int warps : {4, 8, 16}) {
    run<0>("ex2.approx.ftz.f32", warps);
    run<1>("ex2.approx.ftz.bf16x2", warps);
    run<2>("ex2.approx.f16x2", warps);
    run<3>("max.f32 (3 inputs)", warps);
    run<4>("max.f32 + add", warps);
    run<5>("fma.rn.f32x2", warps);
    run<6>("add.rn.f32x2", warps);
    run<7>("fma.rn.f32", warps);
    run<8>("cvt.rn.bf16x2.f32", warps);
  }
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  fence_bench<<<1, 128>>>(out, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("st.shared + fence.proxy.async.shared::cta: %.1f cycles per iteration (4 warps)\n", (double)h / 256);
  return 0;
}
