// allophant_b200 — variable-length self-attention for the wav2vec2 encoder
// (replaces Wav2Vec2Attention's SDPA call, HF:466-549, with a key-padding mask
// derived from per-utterance frame counts instead of a dense [N,1,T,T] mask,
// HF:758-762).
//
// One CTA = 128 query frames of one (utterance, head); head_dim = 64.
//   warps 0..3  softmax: one query row per thread; S read from TMEM, online
//               softmax in fp32, P written bf16 into 128B-swizzled smem, running
//               output O kept in registers (64 fp32) and rescaled per KV block
//   warp 4      TMA producer: Q once, then K_j [128 keys x 64] and Vt_j [64 x 128 keys]
//               through a 2-stage ring
//   warp 5      TMEM allocator + single-thread tcgen05.mma issuer:
//               S = Q K_j^T (128x128x64), PV = P V_j (128x64x128)
// Keys >= lengths[b] get probability exactly 0; KV blocks past the utterance's
// last valid frame and query tiles that are entirely padding are skipped.
#include "aph_common.cuh"

namespace aph {

constexpr int kAttThreads = 192;
constexpr int kAttQ = 128;    // query rows per CTA
constexpr int kAttKV = 128;   // keys per block
constexpr int kAttD = 64;     // head dim
constexpr int kAttTileBytes = 128 * 64 * 2;  // 16 KB
// 7 tiles + barriers = 114,944 B: two CTAs fit one SM (2 x (114,944 + 1,024 reserved) <= 233,472), so the
// tensor core works on one CTA's MMAs while the other CTA's warps are in the softmax.
constexpr int kAttSmemBytes = kAttTileBytes /*Q*/ + 2 * kAttTileBytes /*K*/ + 2 * kAttTileBytes /*Vt*/ +
                              2 * kAttTileBytes /*P*/ + 256 /*barriers*/;
constexpr uint32_t kAttTmemCols = 256;  // S: [0,128)  PV: [128,192)

struct AttParams {
  __nv_bfloat16* ctx;  // [N*T, heads*64]
  const int* lengths;  // [N]
  int T;
  int heads;
  float* lse2;         // [N*heads, T] log2-domain log-sum-exp of every query row (training) or nullptr
};

__global__ void __launch_bounds__(kAttThreads, 2)
    attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttParams p) {
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * kAttQ;
  int len = p.lengths[b];
  len = len < p.T ? len : p.T;
  if (q0 >= len) return;  // whole tile is padding (uniform per CTA, before any barrier/TMEM use)
  const int n_kv = (len + kAttKV - 1) / kAttKV;

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("aph: attention shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kAttTileBytes;       // 2 stages
  uint8_t* s_v = s_k + 2 * kAttTileBytes;   // 2 stages, each two 8 KB halves (keys 0-63 / 64-127)
  uint8_t* s_p = s_v + 2 * kAttTileBytes;   // two 16 KB halves (keys 0-63 / 64-127)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 2 * kAttTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<kAttTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_pv = tmem_base + 128;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kAttTileBytes);
      tma_load_3d(s_q, &tm_q, q_full, 0, q0, bh);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * kAttTileBytes);
        tma_load_3d(s_k + st * kAttTileBytes, &tm_k, &kv_full[st], 0, j * kAttKV, bh);
        tma_load_3d(s_v + st * kAttTileBytes, &tm_v, &kv_full[st], j * kAttKV, 0, bh);
        tma_load_3d(s_v + st * kAttTileBytes + kAttTileBytes / 2, &tm_v, &kv_full[st], j * kAttKV + 64, 0, bh);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64);
      const uint64_t dq = umma_desc_sw128(smem_u32(s_q));
      const uint64_t dp0 = umma_desc_sw128(smem_u32(s_p));
      const uint64_t dp1 = umma_desc_sw128(smem_u32(s_p + kAttTileBytes));
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(s_k + st * kAttTileBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc_s,
                    k != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dv0 = umma_desc_sw128(smem_u32(s_v + st * kAttTileBytes));
        const uint64_t dv1 = umma_desc_sw128(smem_u32(s_v + st * kAttTileBytes + kAttTileBytes / 2));
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = (k < 4 ? dp0 : dp1) + static_cast<uint64_t>(2 * (k & 3));
          const uint64_t db = (k < 4 ? dv0 : dv1) + static_cast<uint64_t>(2 * (k & 3));
          umma_bf16(tmem_pv, da, db, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
        if (j + 1 < n_kv) issue_s(j + 1);  // S is free: p_full(j) implies the softmax warps finished reading it
      }
    }
  } else {
    // ===================== softmax / output (one query row per thread) =====================
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_run = -INFINITY;
    float l_run = 0.f;
    float o[kAttD];
#pragma unroll
    for (int d = 0; d < kAttD; ++d) o[d] = 0.f;
    uint8_t* p_row = s_p + (r >> 3) * 1024 + (r & 7) * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int key0 = j * kAttKV;
      const bool full_block = key0 + kAttKV <= len;  // no padded keys in this block (CTA-uniform)
      // pass 1: row max (scores are already in the log2 domain: q carries head_dim^-0.5 * log2(e))
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c0 = 0; c0 < kAttKV; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_s + lane_off + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (full_block) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx[i & 3] = fmaxf(mx[i & 3], (key0 + c0 + i < len) ? v[i] : -INFINITY);
        }
      }
      const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
      const float alpha = exp2f(m_run - m_new);
      // pass 2: probabilities -> smem (bf16, swizzled K-major), row sum
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c0 = 0; c0 < kAttKV; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_s + lane_off + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (full_block) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = exp2f(v[i] - m_new);
            ls[i & 3] += v[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = (key0 + c0 + i < len) ? exp2f(v[i] - m_new) : 0.f;
            ls[i & 3] += v[i];
          }
        }
        uint8_t* dst_half = p_row + (c0 >> 6) * kAttTileBytes;
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
      l_run = l_run * alpha + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
      m_run = m_new;
      fence_proxy_async_smem();  // P visible to the tensor core's smem reads
      tc_fence_before();         // our TMEM reads of S are ordered before the next S = Q K^T
      mbar_arrive(p_full);

      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < kAttD; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_pv + lane_off + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c0 + i] = fmaf(o[c0 + i], alpha, v[i]);
      }
      tc_fence_before();
    }

    if (q0 + r < p.T) {
      if (p.lse2 != nullptr) p.lse2[static_cast<long long>(bh) * p.T + q0 + r] = m_run + log2f(l_run);
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = p.ctx + (static_cast<long long>(b) * p.T + q0 + r) * (p.heads * kAttD) + h * kAttD;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 o4;
        o4.x = pack_bf16x2(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
        o4.y = pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
        o4.z = pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
        o4.w = pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        d4[i] = o4;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kAttTmemCols>(tmem_base);
  }
}

}  // namespace aph

extern "C" int aph_attention_bf16(const void* q, const void* k, const void* vt, void* ctx,
                                  const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                                  int32_t t_v, void* stream_) {
  return aph_attention_bf16_lse(q, k, vt, ctx, nullptr, lengths, n_utt, heads, T, t_v, stream_);
}

extern "C" int aph_attention_bf16_lse(const void* q, const void* k, const void* vt, void* ctx, float* lse2,
                                      const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T,
                                      int32_t t_v, void* stream_) {
  using namespace aph;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(q && k && vt && ctx && lengths, "null pointer");
  APH_REQUIRE(n_utt > 0 && heads > 0 && T > 0, "empty problem");
  APH_REQUIRE(t_v >= T && t_v % 8 == 0, "t_v must be >= T and a multiple of 8");
  const uint64_t nh = static_cast<uint64_t>(n_utt) * heads;
  CUtensorMap tm_q, tm_k, tm_v;
  {
    const uint64_t dims[3] = {kAttD, static_cast<uint64_t>(T), nh};
    const uint64_t strides[2] = {kAttD * 2, static_cast<uint64_t>(T) * kAttD * 2};
    const uint32_t box[3] = {kAttD, kAttQ, 1};
    int rc = encode_tmap(&tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(t_v), kAttD, nh};
    const uint64_t strides[2] = {static_cast<uint64_t>(t_v) * 2, static_cast<uint64_t>(t_v) * kAttD * 2};
    const uint32_t box[3] = {64, kAttD, 1};
    int rc = encode_tmap(&tm_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    attr_set = true;
  }
  AttParams p;
  p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.lengths = lengths;
  p.T = T;
  p.heads = heads;
  p.lse2 = lse2;
  dim3 grid(ceil_div(T, kAttQ), static_cast<unsigned>(nh));
  attention_kernel<<<grid, kAttThreads, kAttSmemBytes, stream>>>(tm_q, tm_k, tm_v, p);
  APH_POST_LAUNCH(1);
  return APH_OK;
}
