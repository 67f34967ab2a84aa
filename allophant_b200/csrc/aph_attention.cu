// allophant_b200 — variable-length self-attention for the wav2vec2 encoder
// (replaces Wav2Vec2Attention's SDPA call, HF:466-549, with a key-padding mask
// derived from per-utterance frame counts instead of a dense [N,1,T,T] mask,
// HF:758-762).
//
// One CTA = 128 query frames of one (utterance, head); head_dim = 64; keys in blocks of 64;
// two CTAs per SM.  With head_dim 64 the kernel is bound by the softmax, not by the tensor core
// (one exp2 per score on the MUFU pipe, 16/clk/SM, and one TMEM read per score), so the design
// keeps the four softmax warps of a CTA busy all the time:
//   warps 0..3  softmax: one query row per thread; the 64 scores of a block are read from TMEM once
//               and stay in registers (max, exp2, row sum, bf16 pack -> TMEM)
//   warp 4      TMA producer: Q once, then K_j and V_j [64 keys x 64] through two independent 3-stage
//               rings (a K tile is released as soon as its scores are issued); V stays row-major and is
//               read by O += P V as an MN-major B operand, so no transposed copy of V exists anywhere
//   warp 5      TMEM allocator + single-thread tcgen05.mma issuer
//   * scores run TWO blocks ahead: S is double-buffered in TMEM and S_{j+2} = Q K_{j+2}^T is issued the
//     moment the softmax warps have read S_j, so its latency hides behind the softmax of block j+1;
//   * the running output O stays in TMEM and is accumulated by the tensor core across key blocks
//     (O += P_j V_j); it is rescaled only when a row's maximum grows by more than 2^8 (lazy rescaling:
//     P is then expressed relative to a slightly stale maximum, which the final division by the row
//     sum cancels exactly), so nobody waits for the PV product on the common path;
//   * P never touches shared memory: the softmax warps write it (bf16 pairs, tcgen05.st) into the first 32 columns of S_j's own
//     TMEM buffer and O += P V reads it from there as the A operand (the TS form of tcgen05.mma): 32 KB less shared-memory
//     traffic per block and no proxy fence; S_{j+2}, the next writer of that buffer, is issued behind PV_j by the same thread.
// Keys >= lengths[b] get probability exactly 0; key blocks past the utterance's last valid frame and
// query tiles that are entirely padding are skipped.
#include "aph_common.cuh"

namespace aph {

constexpr int kAttThreads = 192;
constexpr int kAttQ = 128;    // query rows per CTA
constexpr int kAttKV = 64;    // keys per block
constexpr int kAttD = 64;     // head dim
constexpr int kAttStages = 3;
constexpr int kAttQBytes = kAttQ * kAttD * 2;    // 16 KB
constexpr int kAttKVBytes = kAttKV * kAttD * 2;  // 8 KB
constexpr int kAttSmemBytes = kAttQBytes + kAttStages * kAttKVBytes /*K*/ + kAttStages * kAttKVBytes /*V*/ + 256 /*barriers*/;
constexpr uint32_t kAttTmemCols = 256;  // S0: [0,64)  S1: [64,128)  O: [128,192)
constexpr float kAttRescaleThreshold = 8.0f;  // log2 units

struct AttParams {
  __nv_bfloat16* ctx;  // [N*T, heads*64]
  const int* lengths;  // [N]
  int T;
  int heads;
  float* lse2;         // [N*heads, T] log2-domain log-sum-exp of every query row (training) or nullptr
  // train-mode attention dropout (HF Wav2Vec2Attention: dropout of the softmax probabilities): element (b*heads+h, q, k)
  // is kept iff the 16-bit half (k & 1) of drop_hash(drop_row_key(seed, (b*heads+h)*T + q), k >> 1) >= threshold
  uint32_t drop_threshold;
  uint32_t drop_seed;
  float drop_scale;
};

// Debug timeline (device buffer set through aph_debug_set_timeline; NULL in production): clock64 stamps of CTA (0,0)
__device__ long long* g_timeline = nullptr;
#define APH_STAMP(slot)                                                                       \
  do {                                                                                        \
    if (g_timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0) g_timeline[slot] = clock64(); \
  } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kDrop>
__global__ void __launch_bounds__(kAttThreads, 2)
    attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttParams p) {
  pdl_trigger();
  const int bh = blockIdx.y;
  const int b = bh / p.heads;
  const int h = bh - b * p.heads;
  const int q0 = blockIdx.x * kAttQ;
  if (threadIdx.x == 0) APH_STAMP(0);

  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("aph: attention shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kAttQBytes;                 // kAttStages tiles
  uint8_t* s_v = s_k + kAttStages * kAttKVBytes;   // kAttStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + kAttStages * kAttKVBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [3]
  uint64_t* v_empty = bars + 10;  // [3]
  uint64_t* s_full = bars + 13;   // [2]  scores of block j in TMEM buffer j & 1
  uint64_t* p_full = bars + 15;   // [2]  probabilities of block j written (and S_j read, O rescaled), barrier j & 1:
                                  //      a warp may run one block ahead of the slowest one (its scores are already
                                  //      there), so consecutive blocks must not share a barrier
  uint64_t* o_full = bars + 17;   // [2]  PV product of block j accumulated (barrier j & 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&o_full[s], 1);
      mbar_init(&p_full[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<kAttTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  if (threadIdx.x == 0) APH_STAMP(1);
  // Everything above (barriers, TMEM) overlapped the previous kernel's tail; frame counts, Q, K and V are its results.
  pdl_wait();
  int len = p.lengths[b];
  len = len < p.T ? len : p.T;
  const int n_kv = (len + kAttKV - 1) / kAttKV;
  const bool active = q0 < len;  // false: the whole tile is padding (uniform per CTA) -> straight to the teardown

  if (!active) {
    // A tile of padded queries only: its context rows are cleared rather than left as they were — the workspace may be shared
    // with other launch lists (engine.WorkspaceArena), and whatever bit patterns they left there would flow through the
    // output projection into K / V rows that P = 0 multiplies (0 x NaN).
    const int r = static_cast<int>(threadIdx.x);
    if (r < kAttQ && q0 + r < p.T) {
      uint4* dst = reinterpret_cast<uint4*>(p.ctx + (static_cast<long long>(b) * p.T + q0 + r) * (p.heads * kAttD) + h * kAttD);
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  } else if (warp == 4) {
    // ===================== TMA producer =====================
    // Whole warp in the loop, ONE elected lane issues (here and in the MMA warp): addresses and coordinates stay warp-uniform
    // and live in uniform registers; under `if (lane == 0)` the compiler wraps every TMA / tcgen05 instruction in an
    // ELECT + R2UR + BRA.U.ANY waterfall (~20 dependent instructions each, profiles/r02_gemm_epilogue_timeline.md).
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kAttQBytes);
      tma_load_3d(s_q, &tm_q, q_full, 0, q0, bh);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % kAttStages;
      const uint32_t ph = static_cast<uint32_t>(j / kAttStages) & 1u;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], kAttKVBytes);
        tma_load_3d(s_k + st * kAttKVBytes, &tm_k, &k_full[st], 0, j * kAttKV, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], kAttKVBytes);
        tma_load_3d(s_v + st * kAttKVBytes, &tm_v, &v_full[st], 0, j * kAttKV, bh);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64) | kIdescBMnMajor;  // B = V [key rows][64 d]: N contiguous
      const uint64_t dq = umma_desc_sw128(smem_u32(s_q));
      mbar_wait(q_full, 0);
      if (lane == 0) APH_STAMP(2);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T into TMEM buffer j & 1
        const int st = j % kAttStages;
        mbar_wait(&k_full[st], static_cast<uint32_t>(j / kAttStages) & 1u);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(s_k + st * kAttKVBytes));
        const uint32_t tmem_s = tmem_base + static_cast<uint32_t>((j & 1) * 64);
        if (elect_one()) {  // the same lane every time: tcgen05.commit tracks the MMAs of the thread that issues it
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
          umma_commit(&s_full[j & 1]);
          umma_commit(&k_empty[st]);  // the K tile is free once these MMAs have read it
        }
        __syncwarp();
      };
      issue_s(0);
      if (n_kv > 1) issue_s(1);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kAttStages;
        mbar_wait(&p_full[j & 1], static_cast<uint32_t>(j >> 1) & 1u);  // P_j in smem, S_j read, O rescaled if needed
        mbar_wait(&v_full[st], static_cast<uint32_t>(j / kAttStages) & 1u);
        tc_fence_after();
        const uint32_t tmem_p = tmem_base + static_cast<uint32_t>((j & 1) * 64);  // P_j sits in the first columns of S_j's buffer
        const uint64_t dv = umma_desc_mn_sw128(smem_u32(s_v + st * kAttKVBytes), kAttKVBytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 keys per UMMA_K step: +8 TMEM columns of P (two bf16 each), +16 rows (2 KB) of V
            umma_bf16_ts(tmem_o, tmem_p + static_cast<uint32_t>(8 * k), dv + static_cast<uint64_t>(128 * k), idesc_pv, (j | k) != 0 ? 1u : 0u);
          umma_commit(&o_full[j & 1]);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
        // S_{j+2} is issued AFTER PV_j and reuses the TMEM buffer of S_j.  tcgen05 operations of one thread complete
        // in order and a commit tracks everything issued before it, so "S_{j+2} ready" also tells the softmax warps
        // that PV_j has finished reading P buffer j & 1 — the buffer P_{j+2} goes to — without a second wait.
        if (j + 2 < n_kv) issue_s(j + 2);
      }
    }
  } else {
    // ===================== softmax / output (one query row per thread) =====================
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float m_ref = -INFINITY;  // reference maximum of the probabilities currently accumulated in O (log2 domain)
    float l_run = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[j & 1], static_cast<uint32_t>(j >> 1) & 1u);
      tc_fence_after();
      const int key0 = j * kAttKV;
      const uint32_t tmem_s = tmem_base + static_cast<uint32_t>((j & 1) * 64) + lane_off;
      float va[32], vb[32];  // keys 0-31 / 32-63 of the block: two plain register arrays
      auto load_scores = [&]() {
        tmem_ld32(tmem_s, va);
        tmem_ld32(tmem_s + 32u, vb);
        tmem_ld_wait();
        if (key0 + kAttKV > len) {  // padded keys in this block (CTA-uniform)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            va[i] = (key0 + i < len) ? va[i] : -INFINITY;
            vb[i] = (key0 + 32 + i < len) ? vb[i] : -INFINITY;
          }
        }
      };
      load_scores();
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx[i & 1] = fmaxf(mx[i & 1], va[i]);
        mx[2 + (i & 1)] = fmaxf(mx[2 + (i & 1)], vb[i]);
      }
      const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      // lazy rescaling: only when this row's maximum outgrows the reference by more than 2^8
      const bool grow = m_blk > m_ref + kAttRescaleThreshold;  // always true for the first block (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? m_blk : m_ref;
        const float alpha = ex2_approx(m_ref - m_new);  // 1 for rows that keep their reference, 0 for the first block
        if (j > 0) {
          // O <- O * alpha in TMEM; the previous PV product must have landed first (warp-collective ld/st)
          mbar_wait(&o_full[(j - 1) & 1], static_cast<uint32_t>((j - 1) >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < kAttD; c0 += 32) {
            float o[32];
            tmem_ld32(tmem_o + lane_off + static_cast<uint32_t>(c0), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= alpha;
            tmem_st32(tmem_o + lane_off + static_cast<uint32_t>(c0), o);
          }
          tmem_st_wait();
          load_scores();  // re-read instead of keeping 64 scores live across the (rare) correction: no spills
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      // score - reference and the row sum as packed fp32 pairs (add.f32x2): half the issue slots of the two most frequent
      // instructions of this loop next to the exponential itself
      const float2 neg_ref = f2_splat(-m_ref);
      float2 lsa = make_float2(0.f, 0.f), lsb = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 da = f2_add(make_float2(va[2 * i], va[2 * i + 1]), neg_ref);
        const float2 db = f2_add(make_float2(vb[2 * i], vb[2 * i + 1]), neg_ref);
        const float2 pa = make_float2(ex2_approx(da.x), ex2_approx(da.y));
        const float2 pb = make_float2(ex2_approx(db.x), ex2_approx(db.y));
        va[2 * i] = pa.x;
        va[2 * i + 1] = pa.y;
        vb[2 * i] = pb.x;
        vb[2 * i + 1] = pb.y;
        lsa = f2_add(lsa, pa);
        lsb = f2_add(lsb, pb);
      }
      l_run += (lsa.x + lsa.y) + (lsb.x + lsb.y);
      if constexpr (kDrop) {
        // dropout of the (still unnormalised) probabilities: the row sum above is taken before it, the 1/(1-p) scale is
        // folded into the final normalisation
        const uint32_t key = drop_row_key(p.drop_seed, static_cast<uint32_t>(bh) * static_cast<uint32_t>(p.T) + static_cast<uint32_t>(q0 + r));
        const uint32_t pair0 = static_cast<uint32_t>(key0 >> 1);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t ha = drop_hash(key, pair0 + i);
          const uint32_t hb = drop_hash(key, pair0 + 16 + i);
          va[2 * i + 0] = drop_keep(ha, 0, p.drop_threshold) ? va[2 * i + 0] : 0.f;
          va[2 * i + 1] = drop_keep(ha, 1, p.drop_threshold) ? va[2 * i + 1] : 0.f;
          vb[2 * i + 0] = drop_keep(hb, 0, p.drop_threshold) ? vb[2 * i + 0] : 0.f;
          vb[2 * i + 1] = drop_keep(hb, 1, p.drop_threshold) ? vb[2 * i + 1] : 0.f;
        }
      }
      // P_j goes into the first 32 columns of S_j's own TMEM buffer (two bf16 per column: the A operand layout of the TS form).
      // Every thread of the warp has read its S row by now, and S_{j+2} — the next writer of this buffer — is issued after
      // PV_j by the same thread, so the tensor core reads P_j before anything overwrites it.
      {
        uint32_t packed[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          packed[i] = pack_bf16x2(va[2 * i], va[2 * i + 1]);
          packed[16 + i] = pack_bf16x2(vb[2 * i], vb[2 * i + 1]);
        }
        tmem_st32(tmem_s, reinterpret_cast<const float*>(packed));
        tmem_st_wait();
      }
      tc_fence_before();         // our TMEM accesses (S read, P / O written) are ordered before the next MMAs
      mbar_arrive(&p_full[j & 1]);
    }

    mbar_wait(&o_full[(n_kv - 1) & 1], static_cast<uint32_t>((n_kv - 1) >> 1) & 1u);
    if (threadIdx.x == 0) APH_STAMP(24);
    tc_fence_after();
    const bool row_in = q0 + r < p.T;
    if (row_in && p.lse2 != nullptr) p.lse2[static_cast<long long>(bh) * p.T + q0 + r] = m_ref + log2f(l_run);
    const float inv = (kDrop ? p.drop_scale : 1.0f) / l_run;
    __nv_bfloat16* dst = p.ctx + (static_cast<long long>(b) * p.T + q0 + r) * (p.heads * kAttD) + h * kAttD;
#pragma unroll
    for (int c0 = 0; c0 < kAttD; c0 += 32) {
      float o[32];
      tmem_ld32(tmem_o + lane_off + static_cast<uint32_t>(c0), o);  // warp-collective: outside the row guard
      tmem_ld_wait();
      if (row_in) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
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
  }

  if (threadIdx.x == 0) APH_STAMP(25);
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kAttTmemCols>(tmem_base);
  }
  if (threadIdx.x == 160) APH_STAMP(26);
}

}  // namespace aph

extern "C" int aph_debug_set_timeline(int64_t* device_buffer) {
  long long* ptr = reinterpret_cast<long long*>(device_buffer);
  APH_CUDA_CHECK(cudaMemcpyToSymbol(aph::g_timeline, &ptr, sizeof(ptr)));
  return APH_OK;
}

extern "C" int aph_attention_bf16(const void* q, const void* k, const void* v, void* ctx,
                                  const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T, void* stream_) {
  return aph_attention_bf16_lse(q, k, v, ctx, nullptr, lengths, n_utt, heads, T, stream_);
}

extern "C" int aph_attention_bf16_lse(const void* q, const void* k, const void* v, void* ctx, float* lse2,
                                      const int32_t* lengths, int32_t n_utt, int32_t heads, int32_t T, void* stream_) {
  return aph_attention_bf16_dropout(q, k, v, ctx, lse2, lengths, n_utt, heads, T, 0u, 0u, 1.0f, stream_);
}

extern "C" int aph_attention_bf16_dropout(const void* q, const void* k, const void* v, void* ctx, float* lse2, const int32_t* lengths,
                                          int32_t n_utt, int32_t heads, int32_t T, uint32_t drop_threshold, uint32_t drop_seed,
                                          float drop_scale, void* stream_) {
  using namespace aph;
  APH_REQUIRE(drop_threshold < 65536u, "drop_threshold is 16 bits (p < 1)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(q && k && v && ctx && lengths, "null pointer");
  APH_REQUIRE(n_utt > 0 && heads > 0 && T > 0, "empty problem");
  const uint64_t nh = static_cast<uint64_t>(n_utt) * heads;
  CUtensorMap tm_q, tm_k, tm_v;
  {
    const uint64_t dims[3] = {kAttD, static_cast<uint64_t>(T), nh};
    const uint64_t strides[2] = {kAttD * 2, static_cast<uint64_t>(T) * kAttD * 2};
    const uint32_t box_q[3] = {kAttD, kAttQ, 1};
    const uint32_t box_k[3] = {kAttD, kAttKV, 1};
    int rc = encode_tmap(&tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box_q, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box_k, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box_k, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    attr_set = true;
  }
  AttParams p;
  p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.lengths = lengths;
  p.T = T;
  p.heads = heads;
  p.lse2 = lse2;
  dim3 grid(ceil_div(T, kAttQ), static_cast<unsigned>(nh));
  p.drop_threshold = drop_threshold;
  p.drop_seed = drop_seed;
  p.drop_scale = drop_scale;
  if (drop_threshold != 0)
    APH_CUDA_CHECK(launch_pdl(attention_kernel<true>, grid, dim3(kAttThreads), kAttSmemBytes, stream, tm_q, tm_k, tm_v, p));
  else
    APH_CUDA_CHECK(launch_pdl(attention_kernel<false>, grid, dim3(kAttThreads), kAttSmemBytes, stream, tm_q, tm_k, tm_v, p));
  APH_POST_LAUNCH(1);
  return APH_OK;
}
