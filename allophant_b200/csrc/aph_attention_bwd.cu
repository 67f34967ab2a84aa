// allophant_b200 — backward of the variable-length self-attention (autograd of Wav2Vec2Attention's
// SDPA call, HF:466-549, reached from `loss.backward()` at estimator.py:738).
//
// Flash-attention style: the probabilities are never stored; each kernel recomputes
// S = Q K^T on the tensor cores and uses the per-row log-sum-exp kept by the forward pass.
// With P = exp2(S - lse2), dP = dO V^T, delta = rowsum(dO o O) and dS = P o (dP - delta):
//     dV = P^T dO          dK = ln2 * dS^T Qs          dQ = head_dim^-0.5 * dS K
// (Qs is the stored query, pre-scaled by head_dim^-0.5 * log2 e; dQ is the gradient of the
// UNSCALED projection output, which is what the QKV weight-gradient GEMM consumes.)
//
//   attention_delta_kernel    delta[b,h,t] = sum_d dO[b,t,h,d] * O[b,t,h,d]
//   attention_bwd_kv_kernel   one CTA = 128 keys of one (utterance, head), loops over query tiles:
//                             S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (bf16, swizzled smem)
//                             -> dV += P^T dO, dK += dS^T Q.  The second pair of MMAs reads the SAME
//                             Q / dO tiles as MN-major B operands, so no transposed copies exist.
//   attention_bwd_q_kernel    one CTA = 128 queries, loops over key tiles: S, dP -> dS -> dQ += dS K
//                             (K tile re-read as an MN-major B operand).
// Outputs go straight into the bf16 [rows][3*hidden] matrix (dQ | dK | dV) that the QKV
// dgrad / wgrad GEMMs read.  Rows of padded frames are written as zeros.
#include <stdlib.h>

#include "aph_common.cuh"

namespace aph {

constexpr int kBwdThreads = 192;  // 4 compute warps (one tile row per thread) + TMA warp + MMA warp
constexpr int kBwdTile = 128;
constexpr int kBwdD = 64;
constexpr int kBwdTileBytes = kBwdTile * kBwdD * 2;  // 16 KB

// probabilities are recomputed as 2^(s - lse): the hardware approximation (relative error 2^-22) like in the forward kernel;
// exp2f() adds range handling for denormal results that a probability rounded to bf16 never needs
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttBwdParams {
  __nv_bfloat16* dqkv;     // [N*T][3*heads*64]
  const float* lse2;       // [N*heads][T]
  const float* delta;      // [N*heads][T]
  const int* lengths;      // [N]
  int T;
  int heads;
  // attention dropout of the forward pass (aph_attention_bf16_dropout): with keep mask M and scale c = 1/(1-p),
  // O = (c M o P) V, so dV = (c M o P)^T dO and dS = P o (c M o dP - delta); delta = rowsum(dO o O) is unchanged
  uint32_t drop_threshold;
  uint32_t drop_seed;
  float drop_scale;
};

// Debug progress markers (host-mapped memory set through aph_debug_set_progress; NULL in production)
__device__ int* g_progress = nullptr;
#define APH_MARK(slot, value)                                              \
  do {                                                                     \
    if (g_progress != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {     \
      g_progress[slot] = (value);                                          \
      __threadfence_system();                                              \
    }                                                                      \
  } while (0)

#define APH_COUNT(slot)                                                    \
  do {                                                                     \
    if (g_progress != nullptr) {                                           \
      atomicAdd_system(g_progress + (slot), 1);                            \
      __threadfence_system();                                              \
    }                                                                      \
  } while (0)

constexpr uint32_t kIdescS = umma_idesc_bf16(128, 128);                       // both K-major
constexpr uint32_t kIdescAcc = umma_idesc_bf16(128, 64) | kIdescBMnMajor;     // B read MN-major ([k rows][64 d])

__device__ __forceinline__ void zero_rows(__nv_bfloat16* dst_row) {
  uint4* d4 = reinterpret_cast<uint4*>(dst_row);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int i = 0; i < 8; ++i) d4[i] = z;
}

// writes 32 bf16 values (one 32-column chunk of a row) into a [128 rows][64 cols]-halved, 128B-swizzled K-major tile
__device__ __forceinline__ void store_chunk_swizzled(uint8_t* tile_row, int c0, int sw, const float* v) {
  uint8_t* dst_half = tile_row + (c0 >> 6) * kBwdTileBytes;
  const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o4;
    o4.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    o4.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    o4.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    o4.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(dst_half + (((chunk0 + i) ^ sw) << 4)) = o4;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                              const __nv_bfloat16* __restrict__ d_o, long long rows,
                                                              int T, int heads, float* __restrict__ delta) {
  // one warp per frame; a lane covers 8 consecutive channels per step, 8 lanes = one head
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int hidden = heads * kBwdD;
  const long long b = row / T;
  const int t = static_cast<int>(row - b * T);
  for (int c = lane * 8; c < hidden; c += 256) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * hidden + c);
    const uint4 g = *reinterpret_cast<const uint4*>(d_o + row * hidden + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = unpack_bf16x2(aw[e]);
      const float2 y = unpack_bf16x2(gw[e]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0) delta[(b * heads + (c >> 6)) * T + t] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// dK, dV: key-stationary
// smem: K 16K | V 16K | Q 2x16K | dO 2x16K | P^T 32K | dS^T 32K | lse/delta 2x2x512 B | barriers
constexpr int kKvSmemBytes = 2 * kBwdTileBytes + 4 * kBwdTileBytes + 4 * kBwdTileBytes + 2048 + 1024 /*dropout row keys*/ + 256;
constexpr uint32_t kKvTmemCols = 512;  // S^T [0,128) dP^T [128,256) dV [256,320) dK [320,384)

__global__ void __launch_bounds__(kBwdThreads, 1)
    attention_bwd_kv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                            const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                            const AttBwdParams p) {
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int k0 = blockIdx.x * kBwdTile;
  int len = p.lengths[b];
  len = len < p.T ? len : p.T;
  const int hidden3 = 3 * p.heads * kBwdD;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (k0 >= len) {  // every key of this tile is padding: gradients are zero (uniform per CTA, before any barrier)
    if (warp < 4) {
      const int r = warp * 32 + lane;
      if (k0 + r < p.T) {
        __nv_bfloat16* row = p.dqkv + (static_cast<long long>(b) * p.T + k0 + r) * hidden3 + h * kBwdD;
        zero_rows(row + p.heads * kBwdD);
        zero_rows(row + 2 * p.heads * kBwdD);
      }
    }
    return;
  }
  const int n_q = (len + kBwdTile - 1) / kBwdTile;  // query tiles that contain valid frames

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("aph: attention backward shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_k = smem;
  uint8_t* s_v = s_k + kBwdTileBytes;
  uint8_t* s_q = s_v + kBwdTileBytes;          // 2 stages
  uint8_t* s_do = s_q + 2 * kBwdTileBytes;     // 2 stages
  uint8_t* s_pt = s_do + 2 * kBwdTileBytes;    // two 16 KB halves (queries 0-63 / 64-127)
  uint8_t* s_dst = s_pt + 2 * kBwdTileBytes;   // two 16 KB halves
  float* s_lse = reinterpret_cast<float*>(s_dst + 2 * kBwdTileBytes);  // [2][128]
  float* s_delta = s_lse + 2 * kBwdTile;                               // [2][128]
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_delta + 2 * kBwdTile);  // [2][128] dropout row keys of the query tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_key + 2 * kBwdTile);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pt_full = bars + 6;
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pt_full, 128);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<kKvTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_st = tmem_base;
  const uint32_t tmem_dpt = tmem_base + 128;
  const uint32_t tmem_dv = tmem_base + 256;
  const uint32_t tmem_dk = tmem_base + 320;

  if (warp == 4) {
    // ===================== TMA producer =====================
    // (whole warp in the loop, one elected lane issues — see aph_attention.cu: no waterfall around the TMA / MMA instructions)
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * kBwdTileBytes);
        tma_load_3d(s_k, &tm_k, kv_full, 0, k0, bh);
        tma_load_3d(s_v, &tm_v, kv_full, 0, k0, bh);
      }
      __syncwarp();
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[st], 2 * kBwdTileBytes);
          tma_load_3d(s_q + st * kBwdTileBytes, &tm_q, &q_full[st], 0, i * kBwdTile, bh);
          tma_load_3d(s_do + st * kBwdTileBytes, &tm_do, &q_full[st], h * kBwdD, i * kBwdTile, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    {
      const uint64_t dk = umma_desc_sw128(smem_u32(s_k));
      const uint64_t dv = umma_desc_sw128(smem_u32(s_v));
      const uint64_t dpt0 = umma_desc_sw128(smem_u32(s_pt));
      const uint64_t dpt1 = umma_desc_sw128(smem_u32(s_pt + kBwdTileBytes));
      const uint64_t dst0 = umma_desc_sw128(smem_u32(s_dst));
      const uint64_t dst1 = umma_desc_sw128(smem_u32(s_dst + kBwdTileBytes));
      mbar_wait(kv_full, 0);
      auto issue_scores = [&](int i) {
        const int st = i & 1;
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(smem_u32(s_q + st * kBwdTileBytes));
        const uint64_t ddo = umma_desc_sw128(smem_u32(s_do + st * kBwdTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // S^T[key][query] = sum_d K[key][d] Q[query][d]
            umma_bf16(tmem_st, dk + static_cast<uint64_t>(2 * k), dq + static_cast<uint64_t>(2 * k), kIdescS, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dP^T[key][query] = sum_d V[key][d] dO[query][d]
            umma_bf16(tmem_dpt, dv + static_cast<uint64_t>(2 * k), ddo + static_cast<uint64_t>(2 * k), kIdescS, k != 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      issue_scores(0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(pt_full, i & 1);
        tc_fence_after();
        // B operands: the same Q / dO tiles, read MN-major ([128 query rows][64 d]); 16 rows per UMMA_K step
        const uint64_t bq = umma_desc_mn_sw128(smem_u32(s_q + st * kBwdTileBytes), kBwdTileBytes);
        const uint64_t bdo = umma_desc_mn_sw128(smem_u32(s_do + st * kBwdTileBytes), kBwdTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // dV[key][d] += sum_query P^T[key][query] dO[query][d]
            const uint64_t da = (k < 4 ? dpt0 : dpt1) + static_cast<uint64_t>(2 * (k & 3));
            umma_bf16(tmem_dv, da, bdo + static_cast<uint64_t>(128 * k), kIdescAcc, (i | k) != 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // dK[key][d] += sum_query dS^T[key][query] Q[query][d]
            const uint64_t da = (k < 4 ? dst0 : dst1) + static_cast<uint64_t>(2 * (k & 3));
            umma_bf16(tmem_dk, da, bq + static_cast<uint64_t>(128 * k), kIdescAcc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(&q_empty[st]);
          if (i + 1 >= n_q) umma_commit(acc_full);
        }
        __syncwarp();
        if (i + 1 < n_q) issue_scores(i + 1);  // S^T / dP^T are free: pt_full(i) means the compute warps finished reading them
      }
    }
  } else {
    // ===================== compute warps: one key row per thread =====================
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const bool key_ok = k0 + r < len;
    uint8_t* pt_row = s_pt + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* dst_row = s_dst + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    for (int i = 0; i < n_q; ++i) {
      // per-query statistics of this tile (guarded: rows past T read as 0; their Q and dO rows are TMA zero fill)
      float* lse_t = s_lse + (i & 1) * kBwdTile;
      float* delta_t = s_delta + (i & 1) * kBwdTile;
      {
        const int qi = i * kBwdTile + r;
        const bool ok = qi < p.T;
        lse_t[r] = ok ? p.lse2[static_cast<long long>(bh) * p.T + qi] : 0.f;
        delta_t[r] = ok ? p.delta[static_cast<long long>(bh) * p.T + qi] : 0.f;
        if (p.drop_threshold != 0)
          s_key[(i & 1) * kBwdTile + r] = drop_row_key(p.drop_seed, static_cast<uint32_t>(bh) * static_cast<uint32_t>(p.T) + static_cast<uint32_t>(qi));
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four compute warps only
      mbar_wait(s_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < kBwdTile; c0 += 32) {
        float s[32], dp[32];
        tmem_ld32(tmem_st + lane_off + static_cast<uint32_t>(c0), s);
        tmem_ld32(tmem_dpt + lane_off + static_cast<uint32_t>(c0), dp);
        tmem_ld_wait();
        if (key_ok && p.drop_threshold != 0) {
          const uint32_t* key_t = s_key + (i & 1) * kBwdTile + c0;
          const uint32_t pair = static_cast<uint32_t>(k0 + r) >> 1;
          const int odd = (k0 + r) & 1;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pr = ex2_fast(s[j] - lse_t[c0 + j]);
            const float keep = drop_keep(drop_hash(key_t[j], pair), odd, p.drop_threshold) ? p.drop_scale : 0.f;
            s[j] = pr * keep;
            dp[j] = pr * (keep * dp[j] - delta_t[c0 + j]);
          }
        } else if (key_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pr = ex2_fast(s[j] - lse_t[c0 + j]);
            s[j] = pr;
            dp[j] = pr * (dp[j] - delta_t[c0 + j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s[j] = 0.f;
            dp[j] = 0.f;
          }
        }
        store_chunk_swizzled(pt_row, c0, sw, s);
        store_chunk_swizzled(dst_row, c0, sw, dp);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pt_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    {
      // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only rows inside the utterance store
      const bool row_in = k0 + r < p.T;
      __nv_bfloat16* row = p.dqkv + (static_cast<long long>(b) * p.T + k0 + r) * hidden3 + h * kBwdD;
      uint4* dk4 = reinterpret_cast<uint4*>(row + p.heads * kBwdD);
      uint4* dv4 = reinterpret_cast<uint4*>(row + 2 * p.heads * kBwdD);
#pragma unroll
      for (int c0 = 0; c0 < kBwdD; c0 += 32) {
        float a[32], g[32];
        tmem_ld32(tmem_dk + lane_off + static_cast<uint32_t>(c0), a);
        tmem_ld32(tmem_dv + lane_off + static_cast<uint32_t>(c0), g);
        tmem_ld_wait();
        constexpr float kLn2 = 0.6931471805599453f;
        if (row_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o4;
            o4.x = pack_bf16x2(a[8 * j + 0] * kLn2, a[8 * j + 1] * kLn2);
            o4.y = pack_bf16x2(a[8 * j + 2] * kLn2, a[8 * j + 3] * kLn2);
            o4.z = pack_bf16x2(a[8 * j + 4] * kLn2, a[8 * j + 5] * kLn2);
            o4.w = pack_bf16x2(a[8 * j + 6] * kLn2, a[8 * j + 7] * kLn2);
            dk4[(c0 >> 3) + j] = o4;
            o4.x = pack_bf16x2(g[8 * j + 0], g[8 * j + 1]);
            o4.y = pack_bf16x2(g[8 * j + 2], g[8 * j + 3]);
            o4.z = pack_bf16x2(g[8 * j + 4], g[8 * j + 5]);
            o4.w = pack_bf16x2(g[8 * j + 6], g[8 * j + 7]);
            dv4[(c0 >> 3) + j] = o4;
          }
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kKvTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// dQ: query-stationary
// smem: Q 16K | dO 16K | K 2x16K | V 2x16K | dS 32K | barriers
constexpr int kQSmemBytes = 2 * kBwdTileBytes + 4 * kBwdTileBytes + 2 * kBwdTileBytes + 256;
constexpr uint32_t kQTmemCols = 512;  // S [0,128) dP [128,256) dQ [256,320)

__global__ void __launch_bounds__(kBwdThreads, 1)
    attention_bwd_q_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                           const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                           const AttBwdParams p, const float q_grad_scale) {
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * kBwdTile;
  int len = p.lengths[b];
  len = len < p.T ? len : p.T;
  const int hidden3 = 3 * p.heads * kBwdD;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (q0 >= len) {  // padded query tile: the forward pass skipped it, its gradient is zero
    if (warp < 4) {
      const int r = warp * 32 + lane;
      if (q0 + r < p.T) zero_rows(p.dqkv + (static_cast<long long>(b) * p.T + q0 + r) * hidden3 + h * kBwdD);
    }
    return;
  }
  const int n_kv = (len + kBwdTile - 1) / kBwdTile;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("aph: attention backward shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_q = smem;
  uint8_t* s_do = s_q + kBwdTileBytes;
  uint8_t* s_k = s_do + kBwdTileBytes;        // 2 stages
  uint8_t* s_v = s_k + 2 * kBwdTileBytes;     // 2 stages
  uint8_t* s_ds = s_v + 2 * kBwdTileBytes;    // two 16 KB halves (keys 0-63 / 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ds + 2 * kBwdTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(ds_full, 128);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (threadIdx.x == 0) APH_MARK(0, 1);
  if (warp == 5) tmem_alloc<kQTmemCols>(tmem_slot);
  if (threadIdx.x == 160) APH_MARK(1, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) APH_MARK(2, 1);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_dp = tmem_base + 128;
  const uint32_t tmem_dq = tmem_base + 256;

  if (warp == 4) {
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kBwdTileBytes);
        tma_load_3d(s_q, &tm_q, q_full, 0, q0, bh);
        tma_load_3d(s_do, &tm_do, q_full, h * kBwdD, q0, b);
      }
      __syncwarp();
      if (lane == 0) APH_MARK(3, 1);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        if (lane == 0) APH_MARK(3, 2 + j);
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[st], 2 * kBwdTileBytes);
          tma_load_3d(s_k + st * kBwdTileBytes, &tm_k, &kv_full[st], 0, j * kBwdTile, bh);
          tma_load_3d(s_v + st * kBwdTileBytes, &tm_v, &kv_full[st], 0, j * kBwdTile, bh);
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    {
      const uint64_t dq = umma_desc_sw128(smem_u32(s_q));
      const uint64_t ddo = umma_desc_sw128(smem_u32(s_do));
      const uint64_t dds0 = umma_desc_sw128(smem_u32(s_ds));
      const uint64_t dds1 = umma_desc_sw128(smem_u32(s_ds + kBwdTileBytes));
      mbar_wait(q_full, 0);
      if (lane == 0) APH_MARK(4, 1);
      auto issue_scores = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        if (lane == 0) APH_MARK(4, 10 + j);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(s_k + st * kBwdTileBytes));
        const uint64_t dv = umma_desc_sw128(smem_u32(s_v + st * kBwdTileBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // S[query][key]
            umma_bf16(tmem_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), kIdescS, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dP[query][key] = sum_d dO[query][d] V[key][d]
            umma_bf16(tmem_dp, ddo + static_cast<uint64_t>(2 * k), dv + static_cast<uint64_t>(2 * k), kIdescS, k != 0 ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      issue_scores(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(ds_full, j & 1);
        if (lane == 0) APH_MARK(5, 1 + j);
        tc_fence_after();
        const uint64_t bk = umma_desc_mn_sw128(smem_u32(s_k + st * kBwdTileBytes), kBwdTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // dQ[query][d] += sum_key dS[query][key] K[key][d]
            const uint64_t da = (k < 4 ? dds0 : dds1) + static_cast<uint64_t>(2 * (k & 3));
            umma_bf16(tmem_dq, da, bk + static_cast<uint64_t>(128 * k), kIdescAcc, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&kv_empty[st]);
          if (j + 1 >= n_kv) umma_commit(acc_full);
        }
        __syncwarp();
        if (j + 1 < n_kv) issue_scores(j + 1);
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const bool row_in = q0 + r < p.T;
    const float lse = row_in ? p.lse2[static_cast<long long>(bh) * p.T + q0 + r] : 0.f;
    const float dl = row_in ? p.delta[static_cast<long long>(bh) * p.T + q0 + r] : 0.f;
    const uint32_t drop_key = drop_row_key(p.drop_seed, static_cast<uint32_t>(bh) * static_cast<uint32_t>(p.T) + static_cast<uint32_t>(q0 + r));
    uint8_t* ds_row = s_ds + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      if (threadIdx.x == 0) APH_MARK(6, 1 + j);
      tc_fence_after();
      const int key0 = j * kBwdTile;
#pragma unroll 1
      for (int c0 = 0; c0 < kBwdTile; c0 += 32) {
        float s[32], dp[32];
        tmem_ld32(tmem_s + lane_off + static_cast<uint32_t>(c0), s);
        tmem_ld32(tmem_dp + lane_off + static_cast<uint32_t>(c0), dp);
        tmem_ld_wait();
        if (p.drop_threshold != 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t hh = drop_hash(drop_key, static_cast<uint32_t>((key0 + c0) >> 1) + i);
            dp[2 * i + 0] = drop_keep(hh, 0, p.drop_threshold) ? dp[2 * i + 0] * p.drop_scale : 0.f;
            dp[2 * i + 1] = drop_keep(hh, 1, p.drop_threshold) ? dp[2 * i + 1] * p.drop_scale : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float pr = (key0 + c0 + i < len) ? ex2_fast(s[i] - lse) : 0.f;
          dp[i] = pr * (dp[i] - dl);
        }
        store_chunk_swizzled(ds_row, c0, sw, dp);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(acc_full, 0);
    if (threadIdx.x == 0) APH_MARK(7, 1);
    tc_fence_after();
    {
      uint4* dq4 = reinterpret_cast<uint4*>(p.dqkv + (static_cast<long long>(b) * p.T + q0 + r) * hidden3 + h * kBwdD);
#pragma unroll
      for (int c0 = 0; c0 < kBwdD; c0 += 32) {
        float a[32];
        tmem_ld32(tmem_dq + lane_off + static_cast<uint32_t>(c0), a);  // warp-collective: outside the row guard
        tmem_ld_wait();
        if (row_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o4;
            o4.x = pack_bf16x2(a[8 * j + 0] * q_grad_scale, a[8 * j + 1] * q_grad_scale);
            o4.y = pack_bf16x2(a[8 * j + 2] * q_grad_scale, a[8 * j + 3] * q_grad_scale);
            o4.z = pack_bf16x2(a[8 * j + 4] * q_grad_scale, a[8 * j + 5] * q_grad_scale);
            o4.w = pack_bf16x2(a[8 * j + 6] * q_grad_scale, a[8 * j + 7] * q_grad_scale);
            dq4[(c0 >> 3) + j] = o4;
          }
        }
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kQTmemCols>(tmem_base);
  }
}

}  // namespace aph

extern "C" int aph_debug_set_progress(int32_t* host_mapped) {
  APH_CUDA_CHECK(cudaMemcpyToSymbol(aph::g_progress, &host_mapped, sizeof(host_mapped)));
  return APH_OK;
}

extern "C" int aph_attention_backward_bf16(const void* q, const void* k, const void* v, const void* ctx, const void* d_ctx,
                                           const float* lse2, float* delta_scratch, void* dqkv,
                                           const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                                           void* stream_) {
  return aph_attention_backward_bf16_dropout(q, k, v, ctx, d_ctx, lse2, delta_scratch, dqkv, lengths, n_utt, heads, T, 0u, 0u, 1.0f, stream_);
}

extern "C" int aph_attention_backward_bf16_dropout(const void* q, const void* k, const void* v, const void* ctx, const void* d_ctx,
                                                   const float* lse2, float* delta_scratch, void* dqkv, const int32_t* lengths,
                                                   int32_t n_utt, int32_t heads, int32_t T, uint32_t drop_threshold,
                                                   uint32_t drop_seed, float drop_scale, void* stream_) {
  using namespace aph;
  APH_REQUIRE(drop_threshold < 65536u, "drop_threshold is 16 bits (p < 1)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(q && k && v && ctx && d_ctx && lse2 && delta_scratch && dqkv && lengths, "null pointer");
  APH_REQUIRE(n_utt > 0 && heads > 0 && T > 0, "empty problem");
  APH_REQUIRE((heads * kBwdD) % 256 == 0, "hidden size must be a multiple of 256");
  const uint64_t nh = static_cast<uint64_t>(n_utt) * heads;
  const uint64_t hidden = static_cast<uint64_t>(heads) * kBwdD;
  CUtensorMap tm_q, tm_k, tm_v, tm_do;
  {
    const uint64_t dims[3] = {kBwdD, static_cast<uint64_t>(T), nh};
    const uint64_t strides[2] = {kBwdD * 2, static_cast<uint64_t>(T) * kBwdD * 2};
    const uint32_t box[3] = {kBwdD, kBwdTile, 1};
    int rc = encode_tmap(&tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  {
    // dO = gradient of the attention context, [n_utt][T][hidden]: a head is a 64-column slice
    const uint64_t dims[3] = {hidden, static_cast<uint64_t>(T), static_cast<uint64_t>(n_utt)};
    const uint64_t strides[2] = {hidden * 2, static_cast<uint64_t>(T) * hidden * 2};
    const uint32_t box[3] = {kBwdD, kBwdTile, 1};
    int rc = encode_tmap(&tm_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d_ctx, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKvSmemBytes));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQSmemBytes));
    attr_set = true;
  }
  const long long rows = static_cast<long long>(n_utt) * T;
  // APH_ATT_BWD_MASK (debug): bit 0 = delta, bit 1 = dK/dV kernel, bit 2 = dQ kernel
  static const int mask = [] {
    const char* e = getenv("APH_ATT_BWD_MASK");
    return e ? atoi(e) : 7;
  }();
  if (mask & 1)
  attention_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(ctx), static_cast<const __nv_bfloat16*>(d_ctx), rows, T, heads, delta_scratch);
  AttBwdParams p;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.lse2 = lse2;
  p.delta = delta_scratch;
  p.lengths = lengths;
  p.T = T;
  p.heads = heads;
  p.drop_threshold = drop_threshold;
  p.drop_seed = drop_seed;
  p.drop_scale = drop_scale;
  dim3 grid(ceil_div(T, kBwdTile), static_cast<unsigned>(nh));
  if (mask & 2) attention_bwd_kv_kernel<<<grid, kBwdThreads, kKvSmemBytes, stream>>>(tm_q, tm_k, tm_v, tm_do, p);
  if (mask & 4) attention_bwd_q_kernel<<<grid, kBwdThreads, kQSmemBytes, stream>>>(tm_q, tm_k, tm_v, tm_do, p, 0.125f);
  APH_POST_LAUNCH(3);
  return APH_OK;
}
