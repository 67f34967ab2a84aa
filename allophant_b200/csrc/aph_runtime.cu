// allophant_b200 — library-level plumbing: error reporting, launch counter,
// run-time resolution of the one driver entry point the kernels need.
#include <atomic>
#include <mutex>
#include <string.h>

#include <stdlib.h>

#include "aph_common.cuh"

namespace aph {

static thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_last_error(const char* what, const char* detail, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s:%d)", what, detail ? detail : "", file,
           line);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled", "driver entry point unavailable (no CUDA driver?)",
                   __FILE__, __LINE__);
    return APH_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "CUresult %d rank %u dims [%llu,%llu,%llu] strides [%llu,%llu] box [%u,%u,%u] base %p",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
             (unsigned long long)(rank > 2 ? gstr[1] : 0), bx[0], rank > 1 ? bx[1] : 0,
             rank > 2 ? bx[2] : 0, base);
    set_last_error("cuTensorMapEncodeTiled", buf, __FILE__, __LINE__);
    return APH_ERR_CUDA;
  }
  return APH_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// -1 = not decided yet (first use reads APH_PDL from the environment; default on)
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("APH_PDL");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

// same scheme for the GEMM's tail split (APH_GEMM_TAIL_SPLIT; default on)
static std::atomic<int> g_tail_split{-1};
bool tail_split_enabled() {
  int v = g_tail_split.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("APH_GEMM_TAIL_SPLIT");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
    g_tail_split.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

}  // namespace aph

extern "C" {

int aph_set_gemm_tail_split(int enabled) {
  const int before = aph::tail_split_enabled() ? 1 : 0;
  aph::g_tail_split.store(enabled ? 1 : 0);
  return before;
}

int aph_set_pdl(int enabled) {
  const int before = aph::pdl_enabled() ? 1 : 0;
  aph::g_pdl.store(enabled ? 1 : 0);
  return before;
}

int aph_abi_version(void) { return APH_ABI_VERSION; }
const char* aph_last_error(void) { return aph::g_last_error; }
int64_t aph_launch_count(void) { return aph::g_launches.load(); }
void aph_reset_launch_count(void) { aph::g_launches.store(0); }

}  // extern "C"
