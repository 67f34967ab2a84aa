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

// 0 = by problem size, 1 = the 64-key kernel, 2 = the persistent pair kernel (aph_set_attention_kernel, APH_ATT_V1=1 / 0)
static std::atomic<int> g_attention_mode{0};

// Debug timeline (device buffer set through aph_debug_set_timeline; NULL in production): clock64 stamps of CTA (0,0)
__device__ long long* g_timeline = nullptr;
#define APH_STAMP(slot)                                                                       \
  do {                                                                                        \
    if (g_timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0) g_timeline[slot] = clock64(); \
  } while (0)

// Steady-state detail of the pair kernel (compiled in with -DAPH_ATT_TIMELINE only): CTA 0, its third item
#ifdef APH_ATT_TIMELINE
#define APH_DSTAMP(cond, slot)                                      \
  do {                                                              \
    if (dstamp_buffer != nullptr && (cond)) dstamp_buffer[slot] = clock64(); \
  } while (0)
#else
#define APH_DSTAMP(cond, slot) \
  do {                         \
  } while (0)
#endif

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


// ================================================================================================================
// Query-tile PAIR kernel (the default): one persistent CTA per SM, 352 threads, works through (utterance, head, pair of
// 128-row query tiles) items.  What the 64-key kernel above cannot hide is the fixed cost per key block of a softmax warp
// (TMEM read latency, P store + wait, barrier round trips: ~450 cycles next to ~512 cycles of MUFU work), plus ~5 k cycles
// of CTA prologue / epilogue per 12 k cycles of blocks (profiles/r02_attention_experiments.md).  Here
//   * keys go in blocks of 128: every barrier round trip and TMEM latency is paid once per 128 exponentials of a thread;
//   * warps 0-3 own query tile A, warps 4-7 tile B (one row per thread), two warps per scheduler share its MUFU pipe;
//   * P has its own TMEM columns, so a tile's score buffer is free again as soon as its warps hold the scores in registers
//     (`s_read`): S_t(j+1) = Q_t K_{j+1}^T runs on the tensor core while block j is exponentiated, and O_t += P_t V_j is
//     issued behind it when P_t(j) arrives — the softmax warps never wait for a tensor-core round trip;
//   * each tile has its own MMA issuer warp (a commit tracks the issuing thread's MMAs): the two tiles' barrier chains do
//     not wait for each other;
//   * K_j / V_j tiles in shared memory serve both query tiles (half the shared-memory fill and L2 traffic per score);
//   * the CTA is persistent: Q of the next item (double-buffered) and its K / V blocks (3-stage rings that run across
//     items) are loaded while the current one computes, TMEM and barriers are set up once.
// TMEM (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512) (bf16 pairs).
constexpr int kPairThreads = 352;
constexpr int kPairKV = 128;
constexpr int kPairStages = 3;
constexpr int kPairQBytes = 2 * kAttQ * kAttD * 2;  // 32 KB: both query tiles, one TMA box of 256 rows
constexpr int kPairKVBytes = kPairKV * kAttD * 2;   // 16 KB
constexpr int kPairSmemBytes = 2 * kPairQBytes + 2 * kPairStages * kPairKVBytes + 2 * kAttQ * kAttD * 2 /*output staging*/ + 256;
constexpr uint32_t kPairTmemCols = 512;
#ifndef APH_ATT_POLY_EVERY
#define APH_ATT_POLY_EVERY 0
#endif
// Lazy rescaling threshold of the pair kernel (log2 units).  P is bf16 and O / the row sums are fp32, so probabilities up to 2^32
// relative to a stale reference are as exact as those below 1 (the 2^8 of the 64-key kernel is the fp16 bound of the published
// kernels); the rescale itself is expensive here — it has to wait for the previous block's PV product, which is issued at the
// end of that block — and with unit-variance q and k (score deviation 8) it fired in most blocks: 90.6 us against 63.8 us.
constexpr float kPairRescaleThreshold = 32.0f;
constexpr int kPairSkewCycles = 3000;  // head start of tile A over tile B at the first item of a CTA (APH_ATT_SKEW overrides)
constexpr int kPolyEvery = APH_ATT_POLY_EVERY;  // one exponential pair in this many on the FMA pipe; 0 = none (4: 63.8 -> 67.3 us, profiles/r02_attention_experiments.md)

// 2^x for a pair of scores on the FMA / ALU pipes instead of the MUFU pipe (Cody-Waite: x = n + f with |f| <= 1/2 through the
// 1.5 * 2^23 rounding constant, a degree-3 fit of 2^f with 7.5e-5 relative error — P is rounded to bf16 afterwards — and n added
// into the exponent field).  Valid for -125 <= x < 128; anything below gives exactly 0 (masked keys are -inf).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float2 xc = make_float2(fmaxf(x.x, -125.f), fmaxf(x.y, -125.f));
  const float2 t = f2_add(xc, f2_splat(12582912.f));
  const float2 n = f2_add(t, f2_splat(-12582912.f));
  const float2 f = f2_fma(n, f2_splat(-1.f), xc);
  float2 q = f2_fma(f2_splat(0.05517165f), f, f2_splat(0.24261113f));
  q = f2_fma(q, f, f2_splat(0.69326097f));
  q = f2_fma(q, f, f2_splat(0.99992806f));
  float2 y;
  y.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  y.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  y.x = x.x < -125.f ? 0.f : y.x;
  y.y = x.y < -125.f ? 0.f : y.y;
  return y;
}

template <bool kDrop>
__global__ void __launch_bounds__(kPairThreads, 1)
    attention_pair_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                          const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const AttParams p,
                          const int n_pairs, const int n_items, const int skew_cycles) {
  pdl_trigger();
  if (threadIdx.x == 0) APH_STAMP(0);
#ifdef APH_ATT_TIMELINE
  long long* const dstamp_buffer = blockIdx.x == 0 ? g_timeline : nullptr;
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("aph: attention shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* s_q = smem;                               // 2 buffers x (tile A, tile B)
  uint8_t* s_k = s_q + 2 * kPairQBytes;              // kPairStages tiles of 128 keys
  uint8_t* s_v = s_k + kPairStages * kPairKVBytes;
  uint8_t* s_o = s_v + kPairStages * kPairKVBytes;   // one output staging tile per query tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_o + 2 * kAttQ * kAttD * 2);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2]  both tiles' last scores of the item that used the buffer are done (2 commits)
  uint64_t* k_full = bars + 4;    // [3]
  uint64_t* k_empty = bars + 7;   // [3]  2 commits: one per tile
  uint64_t* v_full = bars + 10;   // [3]
  uint64_t* v_empty = bars + 13;  // [3]  2 commits: one per tile
  uint64_t* s_full = bars + 16;   // [2]  per query tile: scores of its next block are in TMEM
  uint64_t* s_read = bars + 18;   // [2]  per query tile: those scores are in registers (128 arrivals)
  uint64_t* p_full = bars + 20;   // [2]  per query tile: probabilities written, O rescaled if needed (128 arrivals)
  uint64_t* pv_done = bars + 22;  // [2]  per query tile: O_t += P_t V_j accumulated (P buffer free, O current)
  uint64_t* o_free = bars + 24;   // [2]  per query tile: the item's output has left TMEM (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_read[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
      mbar_init(&o_free[s], 128);
    }
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<kPairTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) APH_STAMP(1);
  pdl_wait();

  if (warp == 8) {
    // ===================== TMA producer =====================
    uint32_t it = 0, g = 0;  // items / key blocks loaded so far (ring positions run across items)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int bh = item / n_pairs;
      const int q0 = (item - bh * n_pairs) * (2 * kAttQ);
      int len = p.lengths[bh / p.heads];
      len = len < p.T ? len : p.T;
      if (q0 >= len) continue;
      const int n_kv = (len + kPairKV - 1) / kPairKV;
      const uint32_t buf = it & 1u;
      mbar_wait(&q_empty[buf], ((it >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[buf], kPairQBytes);
        tma_load_3d(s_q + buf * kPairQBytes, &tm_q, &q_full[buf], 0, q0, bh);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j, ++g) {
        const uint32_t st = g % kPairStages;
        const uint32_t ph = (g / kPairStages) & 1u;
        mbar_wait(&k_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], kPairKVBytes);
          tma_load_3d(s_k + st * kPairKVBytes, &tm_k, &k_full[st], 0, j * kPairKV, bh);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], kPairKVBytes);
          tma_load_3d(s_v + st * kPairKVBytes, &tm_v, &v_full[st], 0, j * kPairKV, bh);
        }
        __syncwarp();
      }
      ++it;
    }
  } else if (warp >= 9) {
    // ===================== MMA issuers: warp 9 query tile A, warp 10 query tile B =====================
    const int t = warp - 9;
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kPairKV);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kAttD) | kIdescBMnMajor;  // B = V [key rows][64 d]: N contiguous
    const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(t * 128);
    const uint32_t tmem_o = tmem_base + 256u + static_cast<uint32_t>(t * 64);
    const uint32_t tmem_p = tmem_base + 384u + static_cast<uint32_t>(t * 64);
    uint32_t it = 0, g = 0;
    uint32_t blocks = 0;  // key blocks of this tile so far (phases of s_full / s_read / p_full / pv_done)
    uint32_t items = 0;   // items of this tile so far (phase of o_free)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int bh = item / n_pairs;
      const int q0 = (item - bh * n_pairs) * (2 * kAttQ);
      int len = p.lengths[bh / p.heads];
      len = len < p.T ? len : p.T;
      if (q0 >= len) continue;
      const int n_kv = (len + kPairKV - 1) / kPairKV;
      const bool both = q0 + kAttQ < len;  // false: tile B is padding only
      const uint32_t buf = it & 1u;
      if (t == 1 && !both) {
        // Tile B has nothing to compute, but its issuer still WALKS the item's barriers and gives its share of every release:
        // skipping ahead would let it wait for a phase two ring revolutions away, which an mbarrier parity cannot tell from
        // the phase that has already completed (it would pass the wait and issue MMAs on stale tiles).
        mbar_wait(&q_full[buf], (it >> 1) & 1u);
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t gj = g + static_cast<uint32_t>(j);
          const uint32_t st = gj % kPairStages;
          mbar_wait(&k_full[st], (gj / kPairStages) & 1u);
          if (elect_one()) {
            umma_commit(&k_empty[st]);
            if (j == n_kv - 1) umma_commit(&q_empty[buf]);
          }
          __syncwarp();
          mbar_wait(&v_full[st], (gj / kPairStages) & 1u);
          if (elect_one()) umma_commit(&v_empty[st]);
          __syncwarp();
        }
        g += static_cast<uint32_t>(n_kv);
        ++it;
        continue;
      }
      mbar_wait(&q_full[buf], (it >> 1) & 1u);
      tc_fence_after();
      const uint64_t dq = umma_desc_sw128(smem_u32(s_q + buf * kPairQBytes + t * (kAttQ * kAttD * 2)));
      // S_t(j) = Q_t K_j^T into the tile's score buffer, once the warps hold the previous block's scores in registers
      auto issue_s = [&](int j) {
        const uint32_t gj = g + static_cast<uint32_t>(j);
        const uint32_t st = gj % kPairStages;
        if (blocks > 0) mbar_wait(&s_read[t], (blocks - 1u) & 1u);  // blocks = index of the block these scores belong to
        mbar_wait(&k_full[st], (gj / kPairStages) & 1u);
        tc_fence_after();
        const uint64_t dk = umma_desc_sw128(smem_u32(s_k + st * kPairKVBytes));
        if (elect_one()) {  // the same lane every time: tcgen05.commit tracks the MMAs of the thread that issues it
          int s_steps = 4;
#ifdef APH_ATT_TIMELINE
          if (!kDrop && (p.drop_seed & 2u)) s_steps = 1;  // experiment: what the score products cost (results wrong)
#endif
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < s_steps) umma_bf16(tmem_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[st]);
          if (j == n_kv - 1) umma_commit(&q_empty[buf]);
        }
        __syncwarp();
      };
      if (skew_cycles > 0 && t == 1 && it == 0) {
        // Tile B starts about one key block behind tile A and stays there (both advance at the same rate): the two tiles'
        // item tails — last PV product, output — no longer coincide, and the tile that is busy has the MUFU pipe to itself
        // while the other one is in its tail.
        const long long start = clock64();
        while (clock64() - start < skew_cycles) {
        }
      }
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t gj = g + static_cast<uint32_t>(j);
        const uint32_t st = gj % kPairStages;
        ++blocks;                          // block j of this item is block (blocks - 1) of the tile
        if (j + 1 < n_kv) issue_s(j + 1);  // waits for s_read of block j
        APH_DSTAMP(it == 2 && lane == 0, 96 + j * 8 + t * 4 + 0);
        // keys past the utterance's end have P = 0: a last block with at most 64 valid keys takes half the PV steps
        int k_steps = (len - j * kPairKV > 64) ? 8 : 4;
#ifdef APH_ATT_TIMELINE
        if (!kDrop && (p.drop_seed & 1u)) k_steps = 1;  // experiment: what the PV products cost (results wrong)
#endif
        mbar_wait(&p_full[t], (blocks - 1u) & 1u);  // P_t(j) written, O_t rescaled if needed
        APH_DSTAMP(it == 2 && lane == 0, 96 + j * 8 + t * 4 + 1);
        mbar_wait(&v_full[st], (gj / kPairStages) & 1u);
        if (j == 0) mbar_wait(&o_free[t], (items & 1u) ^ 1u);  // the previous item's output has been read out of O_t
        tc_fence_after();
        const uint64_t dv = umma_desc_mn_sw128(smem_u32(s_v + st * kPairKVBytes), kPairKVBytes);
        if (elect_one()) {
          // 16 keys per UMMA_K step: +8 TMEM columns of P (two bf16 each), +16 rows (2 KB) of V; both step counts unrolled
          // (a counted loop rebuilds the descriptors through R2UR chains between the MMAs)
          if (k_steps == 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(tmem_o, tmem_p + static_cast<uint32_t>(8 * k), dv + static_cast<uint64_t>(128 * k), idesc_pv, (j | k) != 0 ? 1u : 0u);
          } else if (k_steps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem_o, tmem_p + static_cast<uint32_t>(8 * k), dv + static_cast<uint64_t>(128 * k), idesc_pv, (j | k) != 0 ? 1u : 0u);
          } else {
            umma_bf16_ts(tmem_o, tmem_p, dv, idesc_pv, j != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[t]);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
        APH_DSTAMP(it == 2 && lane == 0, 96 + j * 8 + t * 4 + 2);
      }
      g += static_cast<uint32_t>(n_kv);
      ++it;
      ++items;
    }
  } else {
    // ===================== softmax / output: warps 0-3 tile A, warps 4-7 tile B, one query row per thread =====================
    const int t = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(t * 128) + lane_off;
    const uint32_t tmem_o = tmem_base + 256u + static_cast<uint32_t>(t * 64) + lane_off;
    const uint32_t tmem_p = tmem_base + 384u + static_cast<uint32_t>(t * 64) + lane_off;
    uint32_t blocks = 0, items = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int bh = item / n_pairs;
      const int b = bh / p.heads;
      const int h = bh - b * p.heads;
      const int q0 = (item - bh * n_pairs) * (2 * kAttQ) + t * kAttQ;  // first row of this warp group's tile
      int len = p.lengths[b];
      len = len < p.T ? len : p.T;
      __nv_bfloat16* dst = p.ctx + (static_cast<long long>(b) * p.T + q0 + r) * (p.heads * kAttD) + h * kAttD;
      const bool row_in = q0 + r < p.T;
      if (q0 >= len) {
        // A tile of padded queries only: its context rows are cleared rather than left as they were (see the kernel above)
        if (row_in) {
#pragma unroll
          for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }
      const int n_kv = (len + kPairKV - 1) / kPairKV;
      float m_ref = -INFINITY;  // reference maximum of the probabilities currently accumulated in O (log2 domain)
      float l_run = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int key0 = j * kPairKV;
        const bool two = len - key0 > 64;  // false: the second half of the block is padding only (CTA-uniform)
        mbar_wait(&s_full[t], blocks & 1u);
        tc_fence_after();
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 0);
        // The row's 128 scores pass through the registers 64 at a time (registers are allocated in units of four warps: the
        // 11 warps of this CTA get 168 each, not enough for 128 scores next to the rest): two online-softmax steps per block,
        // keys 0-63 and keys 64-127, each with its own lazy-rescaling decision.  The probabilities of the first half wait in
        // registers (32 of them, bf16 pairs); both halves go to the P columns at the end of the block, because the columns are
        // still being read by the previous block's PV product, which then has the whole block to finish.
        float xa[32], xb[32];
        uint32_t p_lo[32], p_hi[32];
        auto load_half = [&](int half) {
          tmem_ld32(tmem_s + static_cast<uint32_t>(64 * half), xa);
          tmem_ld32(tmem_s + static_cast<uint32_t>(64 * half + 32), xb);
          tmem_ld_wait();
          if (key0 + kPairKV > len) {  // padded keys in this block (CTA-uniform)
            const int k0 = key0 + 64 * half;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              xa[i] = (k0 + i < len) ? xa[i] : -INFINITY;
              xb[i] = (k0 + 32 + i < len) ? xb[i] : -INFINITY;
            }
          }
        };
        // p = 2^(score - reference) in place; subtract and row sum as packed fp32 pairs (add.f32x2)
        auto exp_half = [&]() -> float {
          const float2 neg_ref = f2_splat(-m_ref);
          float2 lsa = make_float2(0.f, 0.f), lsb = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 da = f2_add(make_float2(xa[2 * i], xa[2 * i + 1]), neg_ref);
            const float2 db = f2_add(make_float2(xb[2 * i], xb[2 * i + 1]), neg_ref);
#ifdef APH_ATT_TIMELINE
            if (!kDrop && (p.drop_seed & 4u)) {  // experiment: what the exponentials cost (results wrong)
              lsa = f2_add(lsa, da);
              lsb = f2_add(lsb, db);
              continue;
            }
#endif
            // one pair in kPolyEvery leaves the MUFU pipe (the exponentials are MUFU-bound while both tiles' warps are in them)
            const bool poly = kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery - 1);
            const float2 pa = poly ? ex2_poly2(da) : make_float2(ex2_approx(da.x), ex2_approx(da.y));
            const float2 pb = poly ? ex2_poly2(db) : make_float2(ex2_approx(db.x), ex2_approx(db.y));
            xa[2 * i] = pa.x;
            xa[2 * i + 1] = pa.y;
            xb[2 * i] = pb.x;
            xb[2 * i + 1] = pb.y;
            lsa = f2_add(lsa, pa);
            lsb = f2_add(lsb, pb);
          }
          return (lsa.x + lsa.y) + (lsb.x + lsb.y);
        };
        // One online-softmax step over the 64 scores in xa / xb.  `release`: these are the block's last scores, the score buffer
        // is handed back once the (rare) re-read below can no longer happen.
        auto softmax_half = [&](int half, uint32_t* packed, bool release) {
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            mx0 = fmax3(mx0, xa[2 * i], xa[2 * i + 1]);
            mx1 = fmax3(mx1, xb[2 * i], xb[2 * i + 1]);
          }
          const float m_half = fmaxf(mx0, mx1);
          // lazy rescaling: only when this row's maximum outgrows the reference by more than 2^32
          const bool grow = m_half > m_ref + kPairRescaleThreshold;  // always true for an item's first scores (m_ref = -inf)
          if (__any_sync(0xffffffffu, grow)) {
            const float m_new = grow ? m_half : m_ref;
            const float alpha = ex2_approx(m_ref - m_new);  // 1 for rows that keep their reference, 0 for the first scores
            l_run *= alpha;
            m_ref = m_new;
            if (half == 1) {  // the first half's probabilities of this block are relative to the old reference
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float2 v = unpack_bf16x2(p_lo[i]);
                p_lo[i] = pack_bf16x2(v.x * alpha, v.y * alpha);
              }
            }
            if (j > 0) {
              // O_t <- O_t * alpha in TMEM; the previous PV product must have landed first (warp-collective ld/st).  The scores
              // are read again afterwards instead of staying live across this.
              mbar_wait(&pv_done[t], (blocks - 1u) & 1u);
              tc_fence_after();
#pragma unroll
              for (int c0 = 0; c0 < kAttD; c0 += 32) {
                float o[32];
                tmem_ld32(tmem_o + static_cast<uint32_t>(c0), o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] *= alpha;
                tmem_st32(tmem_o + static_cast<uint32_t>(c0), o);
              }
              tmem_st_wait();
              load_half(half);
            }
          }
          if (release) {
            tc_fence_before();
            mbar_arrive(&s_read[t]);  // every score of the block has been read: the buffer may take the next block
            APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 1);
          }
          // exponentials, dropout of the (still unnormalised) probabilities — the row sum is taken before it, the 1/(1-p) scale
          // is folded into the final normalisation — and the bf16 pairs (the A operand layout of the TS form)
          l_run += exp_half();
          if constexpr (kDrop) {
            const uint32_t key = drop_row_key(p.drop_seed, static_cast<uint32_t>(bh) * static_cast<uint32_t>(p.T) + static_cast<uint32_t>(q0 + r));
            const uint32_t pair0 = static_cast<uint32_t>((key0 >> 1) + 32 * half);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t h0 = drop_hash(key, pair0 + i);
              const uint32_t h1 = drop_hash(key, pair0 + 16 + i);
              xa[2 * i + 0] = drop_keep(h0, 0, p.drop_threshold) ? xa[2 * i + 0] : 0.f;
              xa[2 * i + 1] = drop_keep(h0, 1, p.drop_threshold) ? xa[2 * i + 1] : 0.f;
              xb[2 * i + 0] = drop_keep(h1, 0, p.drop_threshold) ? xb[2 * i + 0] : 0.f;
              xb[2 * i + 1] = drop_keep(h1, 1, p.drop_threshold) ? xb[2 * i + 1] : 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            packed[i] = pack_bf16x2(xa[2 * i], xa[2 * i + 1]);
            packed[16 + i] = pack_bf16x2(xb[2 * i], xb[2 * i + 1]);
          }
        };
        load_half(0);
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 5);
        softmax_half(0, p_lo, !two);
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 2);
        if (two) {
          load_half(1);
          softmax_half(1, p_hi, true);
        }
        // (within an item; the item's first block follows the epilogue's wait for the previous item's last product)
        if (j > 0) mbar_wait(&pv_done[t], (blocks - 1u) & 1u);
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 7);
        tmem_st32(tmem_p, reinterpret_cast<const float*>(p_lo));
        if (two) tmem_st32(tmem_p + 32u, reinterpret_cast<const float*>(p_hi));
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 3);
        tmem_st_wait();
        tc_fence_before();  // our TMEM accesses (P / O written) are ordered before the next MMAs
        mbar_arrive(&p_full[t]);
        ++blocks;
        APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 32 + t * 32 + j * 8 + 4);
      }

      // ---- output of the item: O_t / row sum -> bf16 context rows; TMEM is handed back before the rows are written
      mbar_wait(&pv_done[t], (blocks - 1u) & 1u);
      APH_DSTAMP(items == 2 && (threadIdx.x & 127) == 0, 160 + t * 4 + 0);
      ++items;
      tc_fence_after();
      float o0[32], o1[32];
      tmem_ld32(tmem_o, o0);  // warp-collective: outside the row guard
      tmem_ld32(tmem_o + 32u, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_free[t]);
      APH_DSTAMP(items == 3 && (threadIdx.x & 127) == 0, 160 + t * 4 + 1);
      if (row_in && p.lse2 != nullptr) p.lse2[static_cast<long long>(bh) * p.T + q0 + r] = m_ref + log2f(l_run);
      // The rows leave through a 128-byte-swizzled staging tile and one TMA store per tile (rows past T are clipped by the
      // tensor map): a thread's own row is 128 contiguous bytes, so direct stores would touch 32 lines per instruction.
      const float inv = (kDrop ? p.drop_scale : 1.0f) / l_run;
      uint8_t* stage = s_o + t * (kAttQ * kAttD * 2);
      if ((threadIdx.x & 127) == 0) bulk_store_wait_read<0>();  // the previous item's store has read the staging tile
      named_barrier_sync(1 + t, 128);
      {
        uint8_t* row = stage + r * 128;
        const int sw = r & 7;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o4;
          o4.x = pack_bf16x2(o0[8 * i + 0] * inv, o0[8 * i + 1] * inv);
          o4.y = pack_bf16x2(o0[8 * i + 2] * inv, o0[8 * i + 3] * inv);
          o4.z = pack_bf16x2(o0[8 * i + 4] * inv, o0[8 * i + 5] * inv);
          o4.w = pack_bf16x2(o0[8 * i + 6] * inv, o0[8 * i + 7] * inv);
          *reinterpret_cast<uint4*>(row + ((i ^ sw) << 4)) = o4;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o4;
          o4.x = pack_bf16x2(o1[8 * i + 0] * inv, o1[8 * i + 1] * inv);
          o4.y = pack_bf16x2(o1[8 * i + 2] * inv, o1[8 * i + 3] * inv);
          o4.z = pack_bf16x2(o1[8 * i + 4] * inv, o1[8 * i + 5] * inv);
          o4.w = pack_bf16x2(o1[8 * i + 6] * inv, o1[8 * i + 7] * inv);
          *reinterpret_cast<uint4*>(row + (((4 + i) ^ sw) << 4)) = o4;
        }
      }
      fence_proxy_async_smem();
      named_barrier_sync(1 + t, 128);
      if ((threadIdx.x & 127) == 0) tma_store_3d(&tm_o, stage, h * kAttD, q0, b);
      APH_DSTAMP(items == 3 && (threadIdx.x & 127) == 0, 160 + t * 4 + 2);
    }
    if ((threadIdx.x & 127) == 0) bulk_store_wait_all();  // the staging tiles live in this CTA's shared memory
  }

  if (threadIdx.x == 0) APH_STAMP(25);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<kPairTmemCols>(tmem_base);
  }
  if (threadIdx.x == 0) APH_STAMP(26);
}

}  // namespace aph

extern "C" int aph_set_attention_kernel(int mode) {
  const int before = aph::g_attention_mode.load();
  if (mode >= 0 && mode <= 2) aph::g_attention_mode.store(mode);
  return before;
}

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
  CUtensorMap tm_q, tm_k, tm_v, tm_q2, tm_k2, tm_v2, tm_o;
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
    // the pair kernel: both query tiles in one box of 256 rows, keys in blocks of 128
    const uint32_t box_q2[3] = {kAttD, 2 * kAttQ, 1};
    const uint32_t box_k2[3] = {kAttD, kPairKV, 1};
    rc = encode_tmap(&tm_q2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q, dims, strides, box_q2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_k2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, k, dims, strides, box_k2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    rc = encode_tmap(&tm_v2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v, dims, strides, box_k2, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
    // context rows [n_utt][T][heads * 64]: one 128-row x 64-column tile per query tile, rows past T clipped
    const uint64_t dims_o[3] = {static_cast<uint64_t>(heads) * kAttD, static_cast<uint64_t>(T), static_cast<uint64_t>(n_utt)};
    const uint64_t strides_o[2] = {static_cast<uint64_t>(heads) * kAttD * 2, static_cast<uint64_t>(T) * heads * kAttD * 2};
    const uint32_t box_o[3] = {kAttD, kAttQ, 1};
    rc = encode_tmap(&tm_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ctx, dims_o, strides_o, box_o, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  static int sm_count = 0;
  static int pair_skew = kPairSkewCycles;  // APH_ATT_V1=1: the 64-key, two-CTAs-per-SM kernel (kept for same-box comparisons)
  if (sm_count == 0) {
    int device = 0;
    APH_CUDA_CHECK(cudaGetDevice(&device));
    int count = 0;
    APH_CUDA_CHECK(cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    APH_CUDA_CHECK(cudaFuncSetAttribute(attention_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    const char* v1 = getenv("APH_ATT_V1");
    if (v1 != nullptr && (v1[0] == '0' || v1[0] == '1') && g_attention_mode.load() == 0) g_attention_mode.store(v1[0] == '1' ? 1 : 2);
    const char* skew = getenv("APH_ATT_SKEW");
    if (skew != nullptr) pair_skew = atoi(skew);
    sm_count = count;
  }
  AttParams p;
  p.ctx = static_cast<__nv_bfloat16*>(ctx);
  p.lengths = lengths;
  p.T = T;
  p.heads = heads;
  p.lse2 = lse2;
  p.drop_threshold = drop_threshold;
  p.drop_seed = drop_seed;
  p.drop_scale = drop_scale;
  // 64-key blocks with two CTAs per SM while every persistent CTA would get at most one item (a single utterance: 8.4 us against
  // 11.8), the persistent pair kernel beyond; aph_set_attention_kernel / APH_ATT_V1 force one of them
  const int mode = g_attention_mode.load();
  const uint64_t pair_items = nh * static_cast<uint64_t>(ceil_div(T, 2 * kAttQ));
  const bool use_blocks_of_64 = mode == 1 || (mode == 0 && pair_items <= static_cast<uint64_t>(sm_count));
  if (use_blocks_of_64) {
    dim3 grid(ceil_div(T, kAttQ), static_cast<unsigned>(nh));
    if (drop_threshold != 0)
      APH_CUDA_CHECK(launch_pdl(attention_kernel<true>, grid, dim3(kAttThreads), kAttSmemBytes, stream, tm_q, tm_k, tm_v, p));
    else
      APH_CUDA_CHECK(launch_pdl(attention_kernel<false>, grid, dim3(kAttThreads), kAttSmemBytes, stream, tm_q, tm_k, tm_v, p));
  } else {
    const int n_pairs = ceil_div(T, 2 * kAttQ);
    APH_REQUIRE(nh * static_cast<uint64_t>(n_pairs) < (1ull << 31), "too many (utterance, head, query tile pair) items");
    const int n_items = static_cast<int>(nh) * n_pairs;
    dim3 grid(static_cast<unsigned>(n_items < sm_count ? n_items : sm_count));
    if (drop_threshold != 0)
      APH_CUDA_CHECK(launch_pdl(attention_pair_kernel<true>, grid, dim3(kPairThreads), kPairSmemBytes, stream, tm_q2, tm_k2, tm_v2, tm_o, p, n_pairs, n_items, pair_skew));
    else
      APH_CUDA_CHECK(launch_pdl(attention_pair_kernel<false>, grid, dim3(kPairThreads), kPairSmemBytes, stream, tm_q2, tm_k2, tm_v2, tm_o, p, n_pairs, n_items, pair_skew));
  }
  APH_POST_LAUNCH(1);
  return APH_OK;
}
