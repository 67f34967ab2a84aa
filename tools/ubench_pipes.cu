// Micro-benchmark: issue rate of the instructions of the attention softmax loop on one SM sub-partition
// (MUFU.EX2, FMNMX3, FADD2, F2FP pack), for 1 / 2 / 4 warps per scheduler.  Prints cycles per warp instruction.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/ubench_pipes tools/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kOp>
__global__ void bench(float* out, long long* cycles, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  uint32_t packed = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (kOp == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (kOp == 1) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 1) & 15]), "f"(x[(i + 2) & 15]));
      if (kOp == 2 && (i & 1) == 0) {
        uint64_t a;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[i]), "f"(x[i + 1]));
        asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(a));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(a));
      }
      if (kOp == 3 && (i & 1) == 0) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[i + 1]));
        packed ^= r;
      }
      if (kOp == 4) {  // the loop's mix: per two exponentials one subtract pair, one sum pair, one max3, one pack
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        if ((i & 1) == 0) {
          uint32_t r;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[(i + 4) & 15]), "f"(x[(i + 5) & 15]));
          packed ^= r;
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(packed);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// kVariant 0: max3 + packed subtract + ex2 + packed sum + pack (the kernel's loop); 1: without the maximum; 2: scalar adds;
// 3: ex2 and pack only
template <int kVariant>
__global__ void softmax_loop(const float* in, float* out, long long* cycles, int iters) {
  float acc = 0.f;
  uint32_t pk = 0;
  float m_ref = in[threadIdx.x & 31];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __int_as_float(__float_as_int(m_ref) + i + it);  // cheap distinct inputs
    float mx0 = -1e30f, mx1 = -1e30f;
    float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
    const float2 neg = make_float2(-m_ref, -m_ref);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (kVariant == 0) {
        mx0 = max3(mx0, v[2 * i], v[2 * i + 1]);
        mx1 = max3(mx1, v[32 + 2 * i], v[32 + 2 * i + 1]);
      }
      float2 pa, pb;
      if (kVariant == 2) {
        pa = make_float2(ex2a(v[2 * i] - m_ref), ex2a(v[2 * i + 1] - m_ref));
        pb = make_float2(ex2a(v[32 + 2 * i] - m_ref), ex2a(v[32 + 2 * i + 1] - m_ref));
        ls0.x += pa.x; ls0.y += pa.y; ls1.x += pb.x; ls1.y += pb.y;
      } else if (kVariant == 3) {
        pa = make_float2(ex2a(v[2 * i]), ex2a(v[2 * i + 1]));
        pb = make_float2(ex2a(v[32 + 2 * i]), ex2a(v[32 + 2 * i + 1]));
      } else {
        const float2 da = add2(make_float2(v[2 * i], v[2 * i + 1]), neg);
        const float2 db = add2(make_float2(v[32 + 2 * i], v[32 + 2 * i + 1]), neg);
        pa = make_float2(ex2a(da.x), ex2a(da.y));
        pb = make_float2(ex2a(db.x), ex2a(db.y));
        ls0 = add2(ls0, pa);
        ls1 = add2(ls1, pb);
      }
      pk ^= pack2(pa.x, pa.y) + pack2(pb.x, pb.y);
    }
    acc += ls0.x + ls0.y + ls1.x + ls1.y + mx0 + mx1;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(pk);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  const char* names[5] = {"MUFU.EX2", "FMNMX3", "FADD2 (2 lanes)", "F2FP pack", "EX2 + pack per pair"};
  const int per_iter[5] = {16, 16, 8, 8, 16};
  const int iters = 2000;
  for (int op = 0; op < 5; ++op) {
    for (int warps_per_sched = 1; warps_per_sched <= 4; warps_per_sched *= 2) {
      const int threads = 128 * warps_per_sched;
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) bench<0><<<148, threads>>>(out, cyc, iters);
        if (op == 1) bench<1><<<148, threads>>>(out, cyc, iters);
        if (op == 2) bench<2><<<148, threads>>>(out, cyc, iters);
        if (op == 3) bench<3><<<148, threads>>>(out, cyc, iters);
        if (op == 4) bench<4><<<148, threads>>>(out, cyc, iters);
      }
      long long c = 0;
      cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      const double ops = static_cast<double>(iters) * per_iter[op] * warps_per_sched;
      printf("%-22s %d warp(s) per scheduler: %.2f cycles per warp instruction (scheduler aggregate)\n", names[op], warps_per_sched, c / ops);
    }
  }
  {
    float* in;
    cudaMalloc(&in, 64 * sizeof(float));
    cudaMemset(in, 0, 64 * sizeof(float));
    const char* variants[4] = {"max3 + sub2 + ex2 + sum2 + pack", "sub2 + ex2 + sum2 + pack", "scalar sub + ex2 + scalar sum + pack", "ex2 + pack"};
    for (int variant = 0; variant < 4; ++variant) {
      for (int warps_per_sched = 1; warps_per_sched <= 2; ++warps_per_sched) {
        const int threads = 128 * warps_per_sched;
        for (int rep = 0; rep < 2; ++rep) {
          if (variant == 0) softmax_loop<0><<<148, threads>>>(in, out, cyc, 500);
          if (variant == 1) softmax_loop<1><<<148, threads>>>(in, out, cyc, 500);
          if (variant == 2) softmax_loop<2><<<148, threads>>>(in, out, cyc, 500);
          if (variant == 3) softmax_loop<3><<<148, threads>>>(in, out, cyc, 500);
        }
        long long c = 0;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("softmax loop (%s), %d warp(s) per scheduler: %.0f cycles per 64 scores of a warp, %.2f per exponential (scheduler aggregate)\n",
               variants[variant], warps_per_sched, c / 500.0, c / 500.0 / 64.0 / warps_per_sched);
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
