// Micro-test (B200): may the A operand of tcgen05.mma start at an arbitrary ROW of a larger 128B-swizzled K-major window?
// D_s[128 x 64] = A[s : s + 128, 0:64] * B[64 x 64]^T for row shifts s = 0..16, with the descriptor start address advanced by
// s * 128 bytes and the matrix-descriptor "base offset" field (bits 49-51) either 0 or (start >> 7) & 7.  The positional conv
// re-fetches a 128-row A tile per tap that overlaps the previous one in 127 rows; if a shifted descriptor reads the right rows, one
// window per 64 taps replaces 64 tile loads (profiles/r02_posconv_window.md).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I allophant_b200/csrc -o /tmp/umma_rowshift tools/micro/umma_rowshift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "aph_common.cuh"

using namespace aph;

constexpr int kRows = 160;  // window rows
constexpr int kShifts = 17;

__global__ void __launch_bounds__(128) rowshift_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* out, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                     // kRows x 128 B, SW128 (chunk ^= row & 7)
  uint8_t* s_b = smem + kRows * 128;       // 64 x 128 B (kRows * 128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_b + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kRows * 8; i += 128) {  // 16-byte chunks
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(s_a + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + r * 64 + c * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(s_b + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * 64 + c * 8);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
  for (int s = 0; s < kShifts; ++s) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t start = smem_u32(s_a) + s * 128;
        uint64_t da = umma_desc_sw128(start);
        if (use_base_offset) da |= static_cast<uint64_t>((start >> 7) & 7u) << 49;
        const uint64_t db = umma_desc_sw128(smem_u32(s_b));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, static_cast<uint32_t>(s) & 1u);
    tc_fence_after();
    float v[32];
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    for (int c0 = 0; c0 < 64; c0 += 32) {
      tmem_ld32(tmem + lane_off + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(static_cast<long long>(s) * 128 + tid) * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  std::vector<__nv_bfloat16> ha(kRows * 64), hb(64 * 64);
  std::vector<float> fa(kRows * 64), fb(64 * 64);
  srand(1);
  for (int i = 0; i < kRows * 64; ++i) {
    fa[i] = static_cast<float>(rand() % 17 - 8) / 8.0f;
    ha[i] = __float2bfloat16(fa[i]);
  }
  for (int i = 0; i < 64 * 64; ++i) {
    fb[i] = static_cast<float>(rand() % 13 - 6) / 4.0f;
    hb[i] = __float2bfloat16(fb[i]);
  }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, sizeof(float) * kShifts * 128 * 64);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + 64 * 128 + 64;
  cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(kShifts * 128 * 64);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dout, 0, sizeof(float) * out.size());
    rowshift_kernel<<<1, 128, smem>>>(da, db, dout, mode);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("base_offset mode %d: CUDA error %s\n", mode, cudaGetErrorString(err));
      return 1;
    }
    cudaMemcpy(out.data(), dout, sizeof(float) * out.size(), cudaMemcpyDeviceToHost);
    printf("base_offset %s:", mode ? "(start >> 7) & 7" : "0");
    for (int s = 0; s < kShifts; ++s) {
      double worst = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += static_cast<double>(fa[(s + r) * 64 + k]) * fb[n * 64 + k];
          const double d = fabs(ref - out[(static_cast<long long>(s) * 128 + r) * 64 + n]);
          if (d > worst) worst = d;
        }
      printf(" s=%d:%s", s, worst < 1e-3 ? "ok" : "WRONG");
    }
    printf("\n");
  }
  return 0;
}
