// allophant_b200 — shared device/host helpers for the sm_100a kernels.
//
// Everything here is a thin wrapper over PTX that exists only on Blackwell
// (tcgen05 / TMEM / TMA / mbarrier).  There is no fallback path: the library is
// built for sm_100a only and every launcher returns an error code on failure.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/allophant_b200.h"

namespace aph {

// ---------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------
#define APH_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      aph::set_last_error(#expr, cudaGetErrorString(_e), __FILE__, __LINE__);       \
      return APH_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define APH_REQUIRE(cond, msg)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      aph::set_last_error(#cond, msg, __FILE__, __LINE__);                          \
      return APH_ERR_INVALID;                                                       \
    }                                                                               \
  } while (0)

void set_last_error(const char* what, const char* detail, const char* file, int line);

// Kernel-launch accounting (bench.py's `gpu_launches`).
extern std::atomic<int64_t> g_launches;
#define APH_POST_LAUNCH(n_kernels)            \
  do {                                        \
    aph::g_launches.fetch_add(n_kernels);     \
    APH_CUDA_CHECK(cudaGetLastError());       \
  } while (0)

// cuTensorMapEncodeTiled resolved at run time through the runtime API so the
// library has no link-time dependency on libcuda (it must dlopen on a GPU-less
// host for the symbol-export test).
int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 entries */,
                const uint32_t* box, CUtensorMapSwizzle swizzle);

int sm_count();

// Programmatic dependent launch (PDL): a kernel launched through `launch_pdl` may be scheduled while its predecessor in
// the stream is still draining — its CTAs run their prologue (barrier init, TMEM allocation, descriptor prefetch, constant
// tables into shared memory) and then block in `pdl_wait()` until the predecessor has completed and flushed.  Every kernel
// launched this way MUST call `pdl_wait()` before its first global-memory access; kernels call `pdl_trigger()` at their top
// so that their own successor can start early.  APH_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();
bool tail_split_enabled();  // aph_gemm.cu: the tiles of a partly filled last wave are computed as two half-width items
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

// PDL (see launch_pdl): no-ops when the kernel was launched without the attribute
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (→ launch error reported to the host) instead of hanging the
// device.  The bound is wall-clock (~2 s of SM cycles), far above any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long start = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - start > 4000000000LL) {
      printf("aph: mbarrier wait timed out (block %d,%d thread %d bar %p parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, (void*)bar, parity);
      __trap();
    }
  }
}

// generic-proxy writes (st.shared) → visible to the async proxy (UMMA / TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA -------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 1-D bulk copies (TMA engine, no tensor map): whole rows in and out of shared memory asynchronously
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                   reinterpret_cast<uint64_t>(gmem_dst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// tensor-map (tiled) store smem -> global: rows/columns outside the tensor are clipped by the hardware
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// same, but the tile is ADDED to global memory (element type from the tensor map): split-K partial tiles
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// waits until at most `kPending` of this thread's bulk stores still READ their shared-memory source
template <int kPending>
__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}

// waits until this thread's bulk stores have completed (written, not only read)
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barrier among `threads` threads of the CTA (id 0 is __syncthreads)
__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- thread-block clusters -------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// One CTA of the cluster fetches the box once from L2; the hardware writes it to the same smem
// offset of every CTA in `cta_mask` and completes tx bytes on each of their mbarriers.
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                      int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// ---- CTA-pair (cta_group::2) helpers: two SMs cooperate on one 256-row UMMA ----
// shared::cluster addresses carry the CTA rank; clearing the peer bit addresses the pair's leader
// (rank 0) copy of the same smem offset (cute: Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive (+ expect_tx) on the barrier at this smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint64_t* bar, uint32_t cta, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [remAddr32], %2;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A[both CTAs' smem, 2 x 128 rows] * B[both CTAs' smem, 2 x N/2 rows]; leader CTA issues
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand read from TENSOR MEMORY (the "TS" form: rows = TMEM lanes, K along the columns, two bf16 per 32-bit
// column, 8 columns per UMMA_K = 16 step): the probabilities of the attention kernel go from the softmax warps' registers to the
// tensor core through tcgen05.st without a round trip through shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this
// thread has completed (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// Same, arriving on the barrier at this smem offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row
// (lane base + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// registers -> TMEM, same shape as tmem_ld32 (thread i of the warp writes row lane base + i, 32 columns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (sm_100 UMMA):
//   bits [0,14)  start address >> 4
//   bits [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1)
//   bits [32,46) stride byte offset >> 4  (8 rows x 128 B = 1024 B between row groups)
//   bits [46,48) descriptor version = 1 (Blackwell)
//   bits [49,52) base offset = 0 (tiles are 1024-B aligned)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
// Field layout follows cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// MN-major, 128-byte-swizzled operand (cute::UMMA::make_umma_desc<Major::MN>, canonical layout
// Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): a tile is stored as 64-element
// (128-byte) wide chunks along M/N, each chunk [k rows][128 B]; LBO = bytes between chunks,
// SBO = 1024 B between 8-row groups along K.  One UMMA_K = 16 step advances the start address by 16 rows
// = 2048 B (128 in the encoded address).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((chunk_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t kIdescAMnMajor = 1u << 15;  // cute::UMMA::InstrDescriptor::a_major_
constexpr uint32_t kIdescBMnMajor = 1u << 16;  // cute::UMMA::InstrDescriptor::b_major_
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32, both operands K-major
// (cute::UMMA::InstrDescriptor): c_format(F32)=1 @4, a_format(BF16)=1 @7,
// b_format(BF16)=1 @10, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- math ------------------------------------------------------------------
// Exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) (torch's default, HF "gelu") as
//   y = h + a (1 - e),  h = x / 2,  a = |h|,  e = erfc(|x| / sqrt 2) = 2^R(s),  s = -|x|,
//   R(s) = s (k1 + s (-k2 + s (k3 + s (-k4 + s k5))))
// R is a degree-5 fit of log2 erfc (no constant term: erf(0) = 0 exactly; |erf error| <= 6.4e-7 in fp32 Horner form, the fp32
// round-off class of the GELU; `tests/test_gpu_kernels.py::test_gelu_matches_erf`).  11 FMA-pipe instructions + one MUFU.EX2 per
// element, 6.5 issue slots per element in the packed form below.  Round 1 used Abramowitz-Stegun 7.1.28,
// 1 - (1 + a1 u + ... + a6 u^6)^-16: six Horner steps, four squarings and a reciprocal — 10 issue slots per element packed, and
// the LayerNorm + GELU rows of the feature extractor, the first convolution and the FFN1 epilogue are bound by exactly this
// arithmetic (profiles/r02_gelu.md).  erff() costs ~25 instructions and two MUFU ops.
constexpr float kGeluK1 = 1.15109136f, kGeluK2 = -0.459254674f, kGeluK3 = 0.052561252f, kGeluK4 = 0.00739752032f, kGeluK5 = 0.000520460532f;
__device__ __forceinline__ float gelu_erf(float x) {
  const float s = -fabsf(x);
  float p = fmaf(kGeluK5, s, kGeluK4);
  p = fmaf(p, s, kGeluK3);
  p = fmaf(p, s, kGeluK2);
  p = fmaf(p, s, kGeluK1);
  p *= s;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  const float h = 0.5f * x;
  const float ms = 0.5f * s;           // -a
  const float w = fmaf(s, -0.5f, h);   // h + a
  return fmaf(ms, e, w);               // h + a - a e
}
// Two GELUs at once on the packed fp32 pipe (FFMA2 / FMUL2, sm_100): the same operations in the same order as
// gelu_erf, so each half is bit-identical to the scalar function, at ~10 instead of ~18 issue slots per element.  The
// FFN1 epilogue and the first convolution are bound by exactly this arithmetic.
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  ra = *reinterpret_cast<uint64_t*>(&a);
  rb = *reinterpret_cast<uint64_t*>(&b);
  rc = *reinterpret_cast<uint64_t*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  ra = *reinterpret_cast<uint64_t*>(&a);
  rb = *reinterpret_cast<uint64_t*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  ra = *reinterpret_cast<uint64_t*>(&a);
  rb = *reinterpret_cast<uint64_t*>(&b);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_splat(float v) { return make_float2(v, v); }
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 s = make_float2(-fabsf(x.x), -fabsf(x.y));
  float2 p = f2_fma(f2_splat(kGeluK5), s, f2_splat(kGeluK4));
  p = f2_fma(p, s, f2_splat(kGeluK3));
  p = f2_fma(p, s, f2_splat(kGeluK2));
  p = f2_fma(p, s, f2_splat(kGeluK1));
  p = f2_mul(p, s);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(p.y));
  const float2 h = f2_mul(x, f2_splat(0.5f));
  const float2 ms = f2_mul(s, f2_splat(0.5f));
  const float2 w = f2_fma(s, f2_splat(-0.5f), h);
  return f2_fma(ms, e, w);
}
// Eight pairs at once with the order of operations pinned (volatile asm): Horner step by Horner step ACROSS the pairs, so that
// eight independent dependency chains are in flight.  Left to itself the compiler interleaves only two or three of the sixteen
// chains of a 32-column chunk (register-pressure heuristics), and one warp's epilogue became a serial chain as long as the
// K = 1024 mainloop of its tile (profiles/r02_gemm_epilogue.md).  Same arithmetic as gelu_erf2, bit for bit.
__device__ __forceinline__ void f2v_fma(float2& d, const float2& a, const float2& b, const float2& c) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;"
               : "=l"(*reinterpret_cast<uint64_t*>(&d))
               : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)),
                 "l"(*reinterpret_cast<const uint64_t*>(&c)));
}
__device__ __forceinline__ void f2v_mul(float2& d, const float2& a, const float2& b) {
  asm volatile("mul.rn.f32x2 %0, %1, %2;"
               : "=l"(*reinterpret_cast<uint64_t*>(&d))
               : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
}
__device__ __forceinline__ void gelu_erf2_x8(float2 (&x)[8]) {
  float2 s[8], p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = make_float2(-fabsf(x[i].x), -fabsf(x[i].y));
  const float2 k5 = f2_splat(kGeluK5), k4 = f2_splat(kGeluK4), k3 = f2_splat(kGeluK3), k2 = f2_splat(kGeluK2), k1 = f2_splat(kGeluK1);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(p[i], k5, s[i], k4);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(p[i], p[i], s[i], k3);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(p[i], p[i], s[i], k2);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(p[i], p[i], s[i], k1);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_mul(p[i], p[i], s[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i].x) : "f"(p[i].x));
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p[i].y) : "f"(p[i].y));
  }
  const float2 half = f2_splat(0.5f), minus_half = f2_splat(-0.5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_mul(x[i], x[i], half);              // h
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(x[i], s[i], minus_half, x[i]);  // w = h + a
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_mul(s[i], s[i], half);              // ms = -a
#pragma unroll
  for (int i = 0; i < 8; ++i) f2v_fma(x[i], s[i], p[i], x[i]);        // h + a - a e
}
// d/dx of the GELU above: Phi(x) + x phi(x), Phi from the same erf approximation
// ---- counter-based keep masks for train-mode dropout ----------------------------------------------------------------
// One 32-bit hash per (row, column pair) decides two adjacent elements with 16-bit thresholds: element (row, col) is
// KEPT iff its 16-bit half of drop_hash(drop_row_key(seed, row), col >> 1) is >= threshold (= round(p * 65536)).  The
// backward pass regenerates the mask from (seed, row, col) instead of storing it; tests/helpers.py restates this
// function in numpy to feed the same masks to the oracle.
__host__ __device__ __forceinline__ uint32_t drop_fmix(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_row_key(uint32_t seed, uint32_t row) { return drop_fmix(seed ^ (row * 0x9E3779B1u)); }
__host__ __device__ __forceinline__ uint32_t drop_hash(uint32_t row_key, uint32_t pair) { return drop_fmix(row_key ^ (pair * 0x9E3779B1u)); }
__host__ __device__ __forceinline__ bool drop_keep(uint32_t hash, int odd, uint32_t threshold) {
  return ((odd ? (hash >> 16) : (hash & 0xFFFFu)) >= threshold);
}

__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float u = fabsf(x) * 0.70710678118654752440f;
  float d = fmaf(0.0000430638f, u, 0.0002765672f);
  d = fmaf(d, u, 0.0001520143f);
  d = fmaf(d, u, 0.0092705272f);
  d = fmaf(d, u, 0.0422820123f);
  d = fmaf(d, u, 0.0705230784f);
  d = fmaf(d, u, 1.0f);
  d *= d;
  d *= d;
  d *= d;
  d *= d;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  const float erf_abs = 1.0f - r;                                  // erf(|x| / sqrt 2)
  const float cdf = 0.5f + copysignf(0.5f * erf_abs, x);           // Phi(x)
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);   // phi(x)
  return fmaf(x, pdf, cdf);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace aph
