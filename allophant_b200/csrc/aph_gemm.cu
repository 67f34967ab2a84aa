// allophant_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
// D[m, n] = sum_k A[m, k] * B[n, k]  (A, B bf16 K-major; fp32 accumulators in TMEM)
//
// Roles inside one 320-thread CTA (one CTA per SM, persistent over output tiles):
//   warp 8      TMA producer: A tile 128x64 + HALF of the B tile (BN/2 x 64) per stage, 128B swizzle.
//               CTAs run as PAIRS (cluster of 2, tcgen05 cta_group::2) on a 256 x BN output tile:
//               each CTA owns 128 rows of A and of the accumulator, the B tile is split between
//               the two SMs and read by the pair's single UMMA (M = 256).  Per SM this halves the
//               B bytes held, fetched and read from shared memory per k-block (32 KB instead of
//               48 KB per stage -> 6 stages; operand reads 64 instead of 96 B/clk of the 128 B/clk
//               smem port, which the epilogue's staging traffic otherwise saturates)
//   warp 9      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16).  The two
//               single-thread roles sit on the HIGHEST warp ids: the SMSP arbiter favours high warp
//               ids, so the epilogue's ALU work never delays a TMA or MMA issue
//   warps 0..7  epilogue: tcgen05.ld the accumulator (TMEM lane quadrant = warp % 4, two warps
//               per quadrant split the tile's columns), bias / GELU / residual / padding
//               mask, vectorised global stores
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer tfull/tempty
// (MMA <-> epilogue) so the epilogue of tile i overlaps the mainloop of tile i+1.
//
// The A operand is always fetched through a rank-3 tensor map
// [batch][rows][inner] whose row stride is free, which gives three uses:
//   plain Linear            row stride = K
//   strided Conv1d (HF:281) channels-last input, row stride = conv_stride*C and
//                           inner = kernel*C: output position t reads the
//                           contiguous window starting at sample t*stride — the
//                           rows overlap, no im2col buffer exists
//   grouped pos-conv (HF:326) "taps" mode: k-block j reads rows t - pad + j of
//                           channel group n/64 (zero-filled outside [0, rows))
//
// Training (autograd of the same call sites) reuses the kernel with MN-major operands, so that
// activations, output gradients and weights are read in the layout the forward pass left them in:
//   dgrad  dX = dY W      A = dY K-major, B = W stored [k = out][n = in] (MN-major B)
//   wgrad  dW = dY^T X    A = dY stored [k = frame][m = out], B = X stored [k = frame][n = in]
// An MN-major tile is fetched as 64-column x 64-row boxes (128B swizzle) = the canonical UMMA
// MN-major SW128 layout (LBO = 8 KB between 64-wide chunks, SBO = 1 KB between 8-row groups).
// Frames past the end of a segment are zero-filled by TMA, so K needs no padding.  Split-K work
// items add their partial tile with TMA reduce stores (cp.reduce.async.bulk.tensor ... .add).
#include <string.h>

#include "aph_common.cuh"

namespace aph {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kTapsPerStage = 4;  // sliding-window TAPS mode: B tiles (taps) per ring stage
constexpr int kWinStages = 6;     // ... and ring stages (two 24 KB windows + 6 x 16 KB fit the 160 KB of the 64-column configuration)
constexpr int kGemmThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int kSmemBudget = 200 * 1024;

struct GemmParams {
  int m_tiles_per_batch;
  int batch;
  int n_tiles;
  int k_blocks;
  int a_rows;
  int mode;
  int tap_pad;
  int n;
  int gelu;
  float scale;
  const float* bias;
  const float* resid;
  long long ld_resid;
  float* out_f32;
  long long ld_f32;
  __nv_bfloat16* out_bf16;
  long long ld_bf16;
  long long out_batch_rows;
  const int* lengths;
  int len_period;
  __nv_bfloat16* q;
  __nv_bfloat16* kmat;
  __nv_bfloat16* vt;
  int heads;
  int t_v;
  float q_scale;
  int staged;  // 0 = direct stores, 1 = fp32 output through TMA store, 2 = bf16 output through TMA store
  // ---- MN-major operands / training epilogues
  int a_mn;            // A stored [k][m]
  int b_mn;            // B stored [k][n]
  int k_seq_blocks;    // MN-major: 64-row blocks per segment (k_blocks = segments * k_seq_blocks)
  int b_k_shift;       // MN-major B: row shift inside the segment
  int diag_taps;       // APH_GEMM_DIAG_TAPS: number of taps (0 = off)
  int split_k;         // >= 1; > 1: partial tiles are reduce-added into the (pre-initialised) fp32 output
  int kb_per_split;
  uint32_t idesc;
  __nv_bfloat16* aux_bf16;
  long long ld_aux;
  const __nv_bfloat16* gelu_bwd;
  long long ld_gelu_bwd;
  __nv_bfloat16* vmat;
  uint32_t drop_threshold;
  uint32_t drop_seed;
  float drop_scale;
  int act_bwd;
  // ---- LayerNorm folded into the surrounding GEMMs (inference)
  float2* row_stats;        // producer: [rows][row_stats_slots] (sum, sum of squares) of the stored fp32 values per column slot
  int row_stats_slots;
  const float2* ln_stats;   // consumer: statistics of the A rows, same layout
  int ln_slots;
  const float* ln_colsum;   // consumer: sum_k B[n][k] of the gamma-folded weight
  float ln_inv_cols;        // 1 / (number of columns the statistics run over)
  float ln_eps;
  int taps_span;            // TAPS: input channels one N tile contracts over per tap (64, or a block-diagonal super group)
  // ---- tail split: the tiles of the last, partly filled wave are computed as two 256 x BN/2 halves (see decode_work)
  int total_work;           // work items of the launch (tiles x split_k, or full-wave tiles + 2 x tail tiles)
  int tail_first;           // first work item of the tail (== total_work: no tail split)
  uint32_t idesc_narrow;    // instruction descriptor of the half-width UMMA (N = BN / 2)
  // ---- TAPS with a sliding A window (positional conv): see the producer
  int taps_window;          // taps served by one A window (0: one A tile per tap)
  int win_bytes;            // (taps_window + 128) rows x 128 B
};

// Work item -> (n block, m pair, output batch, k-block range, B row shift)
struct WorkItem {
  int n_blk;
  int m_pair;
  int kb_lo;
  int kb_hi;
  int shift;
  int out_batch;  // DIAG_TAPS: tap index (output batch); otherwise -1
  int narrow;     // 0: whole tile; 1 + h: column half h of the tile (tail split)
};

// Tail split.  With T tiles on C cluster slots the last wave holds R = T mod C tiles; when 2 R <= C each of them becomes
// TWO work items, the left and the right 256 x BN/2 half of the tile (UMMA N = BN/2, half the B rows per stage, two instead of
// four 32-column chunks per epilogue warp).  Every output element still accumulates the same k-blocks in the same order, so
// the result is bit-identical to the unsplit kernel; a half costs ~0.6 of a mainloop (A is read by both) and half an epilogue.
// out-proj / FFN2 of the 32 x 10 s batch have 252-256 tiles on 74 slots (3.4 waves -> 4), the 4.9 k x 1024 GEMMs of a
// training step 80 (1.08 waves -> 2).
__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int work, int m_pairs) {
  WorkItem w;
  w.narrow = 0;
  if (work >= p.tail_first) {
    const int j = work - p.tail_first;
    const int tile = p.tail_first + (j >> 1);
    w.n_blk = tile % p.n_tiles;
    w.m_pair = tile / p.n_tiles;
    w.kb_lo = 0;
    w.kb_hi = p.k_blocks;
    w.shift = p.b_k_shift;
    w.out_batch = -1;
    w.narrow = 1 + (j & 1);
    return w;
  }
  if (p.diag_taps > 0) {
    const int tap = work / m_pairs;
    w.m_pair = work - tap * m_pairs;
    w.n_blk = w.m_pair;
    w.kb_lo = 0;
    w.kb_hi = p.k_blocks;
    w.shift = p.b_k_shift + tap - p.tap_pad;
    w.out_batch = tap;
    return w;
  }
  const int per_split = m_pairs * p.n_tiles;
  const int split = work / per_split;
  const int rest = work - split * per_split;
  w.n_blk = rest % p.n_tiles;
  w.m_pair = rest / p.n_tiles;
  w.kb_lo = split * p.kb_per_split;
  w.kb_hi = w.kb_lo + p.kb_per_split < p.k_blocks ? w.kb_lo + p.kb_per_split : p.k_blocks;
  w.shift = p.b_k_shift;
  w.out_batch = -1;
  return w;
}

// Internal epilogue variant: the generic store epilogue with the fp32 residual fetched by TMA into a second
// per-warp staging tile (one coalesced 4 KB box per 32 x 32 chunk, requested one chunk ahead) instead of one
// 128-byte line per thread: row-strided residual loads cost 24 us of a 61 us out-proj GEMM.
constexpr int kEpiStoreResidTma = 2;
// Same, and the kernel also (a) writes a bf16 copy of the stored values through a third per-warp staging tile + TMA and (b)
// leaves per-row partial sums / sums of squares of the stored values: the LayerNorm that follows in the encoder is then applied
// inside the NEXT GEMM's epilogue (`ln_stats`), which reads the bf16 copy as its A operand — no LayerNorm kernel runs.
constexpr int kEpiStoreResidStats = 3;

template <int BN, int EPI = APH_EPI_STORE>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / 2) * kBK * 2;  // this CTA's half of the pair's B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // one 32-row x 128-byte output staging tile per epilogue warp (+ one residual tile in the TMA-residual variant)
  static constexpr int kEpiBytes = EPI == kEpiStoreResidStats ? 8 * 12288 : (EPI == kEpiStoreResidTma ? 8 * 8192 : 8 * 4096);
  // everything has to fit the 227 KB a CTA can opt into: stages + epilogue staging + alignment slack + barriers
  // per-tile column vectors (bias | folded-LayerNorm column sums), BN floats each; the dynamic shared memory base must be
  // 1024-byte aligned (checked at kernel entry), so no alignment slack is reserved
  static constexpr int kVecBytes = 2048;
  static constexpr int kBudget = 232448 - kEpiBytes - kVecBytes - 512;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // double-buffered accumulator (128 lanes x BN columns per CTA)
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kVecBytes + 512 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
    gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                     const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_resid,
                     const __grid_constant__ CUtensorMap tm_copy, const __grid_constant__ CUtensorMap tm_b_narrow,
                     const __grid_constant__ CUtensorMap tm_a_win, const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) {  // 128B-swizzled tiles need 1024-byte alignment
    if (threadIdx.x == 0) printf("aph: GEMM shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* epi_smem = smem + kStages * Cfg::kStageBytes;  // 1024-byte aligned (stage sizes are multiples of 8 KB)
  float* vec_smem = reinterpret_cast<float*>(epi_smem + Cfg::kEpiBytes);  // [0, 256) bias, [256, 512) LayerNorm column sums
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes + Cfg::kVecBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* resid_bar = tempty_bar + 2;  // [8] one per epilogue warp (TMA-residual variant)
  uint64_t* awin_full = resid_bar + 8;   // [2] A windows of the sliding-window TAPS mode
  uint64_t* awin_empty = awin_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(awin_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_trigger();  // the next kernel in the stream may start its prologue while this one runs
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (p.staged || EPI == APH_EPI_QKV) tma_prefetch_desc(&tm_out);
    if (EPI == APH_EPI_QKV) {
      tma_prefetch_desc(&tm_resid);
      tma_prefetch_desc(&tm_copy);
    }
    if (EPI == kEpiStoreResidTma || EPI == kEpiStoreResidStats) tma_prefetch_desc(&tm_resid);
    if (EPI == kEpiStoreResidStats) tma_prefetch_desc(&tm_copy);
    if (p.tail_first < p.total_work && !p.b_mn) tma_prefetch_desc(&tm_b_narrow);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's copy is the one in use: one arrive.expect_tx per CTA of the pair
      mbar_init(&empty_bar[s], 1);  // released in both CTAs by the leader's tcgen05.commit multicast
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);  // leader's copy: one arrival per epilogue warp of both CTAs
    }
    for (int s = 0; s < 8; ++s) mbar_init(&resid_bar[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&awin_full[s], 2);
      mbar_init(&awin_empty[s], 1);
    }
    if (p.taps_window > 0) tma_prefetch_desc(&tm_a_win);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // barrier inits visible cluster-wide before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  // Work item = (N tile, pair of M tiles); the two CTAs of a cluster take the two M tiles.
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int tiles_m = p.m_tiles_per_batch * p.batch;
  const int m_pairs = (tiles_m + 1) >> 1;
  const int total_work = p.total_work;

  if (warp == 8) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (coordinates and barrier addresses stay warp-uniform, so they live in uniform
    // registers) and ONE elected lane issues; under `if (lane == 0)` every TMA / MMA instruction was wrapped in an
    // ELECT + R2UR + BRA.U.ANY waterfall, ~20 dependent instructions per tcgen05.mma on the issuing thread.
    {
      int stage = 0;
      uint32_t phase = 0;
      int wbuf = 0;
      uint32_t wphase = 0;
      for (int work = cluster_id; work < total_work; work += n_clusters) {
        const WorkItem w = decode_work(p, work, m_pairs);
        const int n_blk = w.n_blk;
        int mt = 2 * w.m_pair + cta_rank;
        if (mt >= tiles_m) mt = tiles_m - 1;  // odd tile count: the idle CTA still feeds the shared B half
        const int b = mt / p.m_tiles_per_batch;
        const int t0 = (mt % p.m_tiles_per_batch) * kBM;
        if (p.taps_window > 0) {
          // Sliding A window (grouped positional conv, one 64-channel group per N tile): tap j reads rows t0 - pad + j .. + 127 of
          // the group's channels, so consecutive taps overlap in 127 of 128 rows.  ONE window of W + 127 rows serves W taps — the
          // MMA issuer moves the A descriptor's start address down one 128-byte row per tap (the 128B swizzle is a function of
          // the absolute shared-memory address, so a row-shifted descriptor reads what TMA wrote: tools/micro/umma_rowshift.cu) —
          // and only the B tile (4 KB per CTA) still moves per tap.  One 16 KB A tile per tap was 160 B/clk/SM of L2 traffic for
          // 128 cycles of tensor work: the kernel ran at 0.3 of the tensor peak (profiles/r02_posconv_window.md).
          const int W = p.taps_window;
          uint8_t* b_ring = smem + 2 * p.win_bytes;
          for (int c = 0; c * W < p.k_blocks; ++c) {
            mbar_wait(&awin_empty[wbuf], wphase ^ 1);
            if (elect_one()) {
              if (cta_rank == 0) {
                mbar_arrive_expect_tx(&awin_full[wbuf], static_cast<uint32_t>(p.win_bytes));
              } else {
                mbar_arrive_expect_tx_remote(&awin_full[wbuf], 0, static_cast<uint32_t>(p.win_bytes));
              }
              tma_load_3d_pair(smem + wbuf * p.win_bytes, &tm_a_win, &awin_full[wbuf], n_blk * kBK, t0 - p.tap_pad + c * W, b);
            }
            __syncwarp();
            wbuf ^= 1;
            if (wbuf == 0) wphase ^= 1;
            for (int jl = 0; jl < W; jl += kTapsPerStage) {  // a ring stage holds the B tiles of kTapsPerStage taps
              const int kb = c * W + jl;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                if (cta_rank == 0) {
                  mbar_arrive_expect_tx(&full_bar[stage], kTapsPerStage * Cfg::kBBytes);
                } else {
                  mbar_arrive_expect_tx_remote(&full_bar[stage], 0, kTapsPerStage * Cfg::kBBytes);
                }
#pragma unroll
                for (int g = 0; g < kTapsPerStage; ++g)
                  tma_load_2d_pair(b_ring + (stage * kTapsPerStage + g) * Cfg::kBBytes, &tm_b, &full_bar[stage], (kb + g) * kBK,
                                   n_blk * BN + cta_rank * (BN / 2));
              }
              __syncwarp();
              if (++stage == kWinStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          continue;
        }
        for (int kb = w.kb_lo; kb < w.kb_hi; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (elect_one()) {
          // both CTAs' bytes are accounted on the LEADER's full barrier (the leader issues the pair's MMA)
          const uint32_t stage_bytes = w.narrow != 0 ? Cfg::kABytes + Cfg::kBBytes / 2 : Cfg::kStageBytes;
          if (cta_rank == 0) {
            mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          } else {
            mbar_arrive_expect_tx_remote(&full_bar[stage], 0, stage_bytes);
          }
          int seg = 0, r0 = kb * kBK;
          if (p.a_mn | p.b_mn) {
            seg = kb / p.k_seq_blocks;
            r0 = (kb - seg * p.k_seq_blocks) * kBK;
          }
          if (p.a_mn) {
            // [64 frames][64 output channels] boxes: two 64-wide chunks cover this CTA's 128 rows of D
            tma_load_3d_pair(sa, &tm_a, &full_bar[stage], mt * kBM, r0, seg);
            tma_load_3d_pair(sa + Cfg::kABytes / 2, &tm_a, &full_bar[stage], mt * kBM + 64, r0, seg);
          } else if (p.mode == APH_GEMM_TAPS) {
            // k-block kb = (tap, 64-channel slice of the span this output tile contracts over)
            const int kpt = p.taps_span / kBK;
            const int tap = kb / kpt;
            const int chan = (n_blk * kBK / p.taps_span) * p.taps_span + (kb - tap * kpt) * kBK;
            tma_load_3d_pair(sa, &tm_a, &full_bar[stage], chan, t0 - p.tap_pad + tap, b);
          } else {
            tma_load_3d_pair(sa, &tm_a, &full_bar[stage], kb * kBK, t0, b);
          }
          if (w.narrow != 0) {  // column half of a tail tile: this CTA feeds BN/4 of its BN/2 B rows
            const int n0 = n_blk * BN + (w.narrow - 1) * (BN / 2) + cta_rank * (BN / 4);
            if (p.b_mn) {
#pragma unroll
              for (int c = 0; c < (BN / 4) / 64; ++c)
                tma_load_3d_pair(sb + c * (64 * kBK * 2), &tm_b, &full_bar[stage], n0 + 64 * c, r0 + w.shift, seg);
            } else {
              tma_load_2d_pair(sb, &tm_b_narrow, &full_bar[stage], kb * kBK, n0);
            }
          } else if (p.b_mn) {
            const int n0 = n_blk * BN + cta_rank * (BN / 2);
#pragma unroll
            for (int c = 0; c < (BN / 2) / 64; ++c)
              tma_load_3d_pair(sb + c * (64 * kBK * 2), &tm_b, &full_bar[stage], n0 + 64 * c, r0 + w.shift, seg);
          } else {
            tma_load_2d_pair(sb, &tm_b, &full_bar[stage], kb * kBK, n_blk * BN + cta_rank * (BN / 2));
          }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (one thread) =====================
    if (cta_rank == 0) {
      const uint32_t idesc = p.idesc;
      // descriptor advance per UMMA_K = 16: 32 bytes along K inside the swizzle atom (K-major) or 16 rows of
      // 128 bytes (MN-major), in 16-byte units
      const uint64_t a_step = p.a_mn ? 128u : 2u;
      const uint64_t b_step = p.b_mn ? 128u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int wbuf = 0;
      uint32_t wphase = 0;
      for (int work = cluster_id; work < total_work; work += n_clusters) {
        const WorkItem w = decode_work(p, work, m_pairs);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        const uint32_t item_idesc = w.narrow != 0 ? p.idesc_narrow : idesc;
        if (p.taps_window > 0) {  // sliding A window: see the producer
          const int W = p.taps_window;
          const uint32_t b_ring = smem_u32(smem + 2 * p.win_bytes);
          for (int c = 0; c * W < p.k_blocks; ++c) {
            mbar_wait(&awin_full[wbuf], wphase);
            tc_fence_after();
            const uint32_t win = smem_u32(smem + wbuf * p.win_bytes);
            for (int jl = 0; jl < W; jl += kTapsPerStage) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint64_t da = umma_desc_sw128(win + static_cast<uint32_t>(jl) * 128u);
              const uint64_t db = umma_desc_sw128(b_ring + static_cast<uint32_t>(stage * kTapsPerStage) * Cfg::kBBytes);
              if (elect_one()) {
#pragma unroll
                for (int g = 0; g < kTapsPerStage; ++g) {  // next tap: A one row (8 x 16 B) down, B one tile on
#pragma unroll
                  for (int k = 0; k < kBK / 16; ++k)
                    umma_bf16_pair(tmem_d, da + static_cast<uint64_t>(8 * g + 2 * k), db + static_cast<uint64_t>(g * (Cfg::kBBytes >> 4) + 2 * k), idesc,
                                   (c | jl | g | k) != 0 ? 1u : 0u);
                }
                umma_commit_pair(&empty_bar[stage], static_cast<uint16_t>(3));
                if (jl + kTapsPerStage >= W) umma_commit_pair(&awin_empty[wbuf], static_cast<uint16_t>(3));  // the window is free in both CTAs
              }
              __syncwarp();
              if (++stage == kWinStages) {
                stage = 0;
                phase ^= 1;
              }
            }
            wbuf ^= 1;
            if (wbuf == 0) wphase ^= 1;
          }
          if (elect_one()) umma_commit_pair(&tfull_bar[acc], static_cast<uint16_t>(3));
          __syncwarp();
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
          continue;
        }
        for (int kb = w.kb_lo; kb < w.kb_hi; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = p.a_mn ? umma_desc_mn_sw128(sa, 64 * kBK * 2) : umma_desc_sw128(sa);
          const uint64_t db = p.b_mn ? umma_desc_mn_sw128(sa + Cfg::kABytes, 64 * kBK * 2) : umma_desc_sw128(sa + Cfg::kABytes);
          if (elect_one()) {  // the same lane every time: tcgen05.commit tracks the MMAs of the thread that issues it
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              umma_bf16_pair(tmem_d, da + static_cast<uint64_t>(k) * a_step, db + static_cast<uint64_t>(k) * b_step, item_idesc,
                             (kb != w.kb_lo || k != 0) ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage], static_cast<uint16_t>(3));  // frees the stage in both CTAs
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_pair(&tfull_bar[acc], static_cast<uint16_t>(3));  // accumulator ready in both CTAs
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (8 warps: 128 rows x 2 column halves) =====================
    constexpr bool kResid = EPI == kEpiStoreResidTma || EPI == kEpiStoreResidStats;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = warp >> 2;          // which half of the tile's columns this warp drains
    const int r = quad * 32 + lane;
    // Output staging: each warp owns a 32-row x 128-byte tile (128B swizzle).  Registers -> smem
    // (conflict-free 16-byte writes) -> one coalesced TMA store per tile; the hardware clips rows and
    // columns outside the output.  Row-strided 16-byte global stores from registers cost ~25 % of
    // the whole GEMM (1.50 vs 1.13 PFLOP/s with the stores removed).
    uint8_t* stage_buf = epi_smem + warp * 4096;
    uint8_t* stage_row = stage_buf + lane * 128;
    const int sw = lane & 7;
    bool store_pending = false;  // a TMA store may still be reading stage_buf
    // TMA-residual variant: second tile per warp, filled by cp.async.bulk.tensor one chunk ahead of its use
    uint8_t* resid_buf = epi_smem + 8 * 4096 + warp * 4096;
    const uint8_t* resid_row = resid_buf + lane * 128;
    uint32_t resid_phase = 0;
    // stats variant: third tile per warp, 32 rows x 64 bf16 columns (two chunks), for the bf16 copy of the output
    uint8_t* copy_buf = epi_smem + 16 * 4096 + warp * 4096;
    uint8_t* copy_row = copy_buf + lane * 128;

    int acc = 0;
    uint32_t acc_phase = 0;
    // Column vectors of the tile (bias, LayerNorm column sums): thread i of the 256 epilogue threads owns column i of the tile,
    // fetches its two values ONE TILE AHEAD into registers and parks them in shared memory at the start of the tile; the chunks
    // then read them as broadcast LDS.  As 8 x LDG.128 per chunk they cost 400-1000 cycles of L2 latency per chunk on the one
    // warp whose 4 chunks in series set the tile period (profiles/r02_gemm_epilogue_timeline.md).
    const bool has_bias = p.bias != nullptr, has_colsum = p.ln_colsum != nullptr;
    // epilogue options read once (kernel parameters live in the constant bank: tested per chunk they cost a constant load and a
    // dependent branch each, ~500 cycles per chunk in the timeline)
    const int act = p.gelu, staged = p.staged;
    const bool has_ln = p.ln_stats != nullptr, has_aux = p.aux_bf16 != nullptr, has_act_bwd = p.gelu_bwd != nullptr;
    const bool has_drop = p.drop_threshold != 0, plain_resid = !kResid && p.resid != nullptr;
    const bool direct_f32 = p.out_f32 != nullptr && staged != 1;
    const bool direct_bf16 = p.out_bf16 != nullptr && staged != 2 && EPI != kEpiStoreResidStats;
    const float out_scale = p.scale;
    auto fetch_vectors = [&](int work_item, float& bias_value, float& colsum_value) {
      bias_value = 0.f;
      colsum_value = 0.f;
      if (work_item < total_work && static_cast<int>(threadIdx.x) < BN) {
        const WorkItem wn = decode_work(p, work_item, m_pairs);
        const int vcol = (wn.out_batch >= 0 ? 0 : wn.n_blk * BN) + static_cast<int>(threadIdx.x);
        if (vcol < p.n) {
          if (has_bias) bias_value = __ldg(p.bias + vcol);
          if (has_colsum) colsum_value = __ldg(p.ln_colsum + vcol);
        }
      }
    };
    float bias_next, colsum_next;
    fetch_vectors(cluster_id, bias_next, colsum_next);
    for (int work = cluster_id; work < total_work; work += n_clusters) {
      const WorkItem w = decode_work(p, work, m_pairs);
      const int n_blk = w.n_blk;
      const int mt = 2 * w.m_pair + cta_rank;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // every epilogue warp is done with the previous tile's vectors
      if (static_cast<int>(threadIdx.x) < BN) {
        vec_smem[threadIdx.x] = bias_next;
        vec_smem[256 + threadIdx.x] = colsum_next;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      fetch_vectors(work + n_clusters, bias_next, colsum_next);  // in flight while this tile drains
      const int b = w.out_batch >= 0 ? w.out_batch : mt / p.m_tiles_per_batch;
      const int t = (mt % p.m_tiles_per_batch) * kBM + r;
      const int col_base = w.out_batch >= 0 ? 0 : n_blk * BN;  // DIAG_TAPS: the 256 columns of the diagonal block
      const bool row_ok = mt < tiles_m && t < p.a_rows;
      const long long grow = static_cast<long long>(b) * p.out_batch_rows + t;
      bool masked = false;
      int utt = 0, tt = 0;
      if (p.len_period > 0) {
        utt = static_cast<int>(grow / p.len_period);
        tt = static_cast<int>(grow - static_cast<long long>(utt) * p.len_period);
        if (p.lengths != nullptr && row_ok) masked = tt >= p.lengths[utt];
      }
      const int t_tile_row0 = (mt % p.m_tiles_per_batch) * kBM + quad * 32;
      const bool resid_tma = kResid && mt < tiles_m;  // warp-uniform
      float row_sum = 0.f, row_sq = 0.f;  // stats variant: this thread's row over this warp's columns
      // LayerNorm of the A rows applied here: y = rstd * (acc - mean * colsum[n]) + bias'[n]
      float ln_rstd = 1.f, ln_nmr = 0.f;
      if (has_ln && row_ok) {
        // all partial sums of the row in flight at once (one L2 round trip, not ln_slots of them): 16-byte loads of two
        // slots each, ln_slots even and <= 16 (checked by the launcher)
        const float4* st = reinterpret_cast<const float4*>(p.ln_stats + grow * p.ln_slots);
        float4 part[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] = 2 * i < p.ln_slots ? __ldg(st + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1 += part[i].x + part[i].z;
          s2 += part[i].y + part[i].w;
        }
        const float mean = s1 * p.ln_inv_cols;
        const float var = fmaxf(s2 * p.ln_inv_cols - mean * mean, 0.f);
        ln_rstd = rsqrtf(var + p.ln_eps);
        ln_nmr = -mean * ln_rstd;
      }
      // columns this warp finishes: its half of the tile, or — tail split — its half of the item's column half, which sits in
      // the first BN/2 accumulator columns
      int c_begin = half * (BN / 2), c_end = c_begin + BN / 2, tmem_shift = 0;
      if (w.narrow != 0) {
        tmem_shift = (w.narrow - 1) * (BN / 2);
        c_begin = tmem_shift + half * (BN / 4);
        c_end = c_begin + BN / 4;
      }
      if (resid_tma && col_base + c_begin < p.n && lane == 0) {  // first chunk: requested before the accumulator is ready
        mbar_arrive_expect_tx(&resid_bar[warp], 4096);
        tma_load_3d(resid_buf, &tm_resid, &resid_bar[warp], col_base + c_begin, t_tile_row0, b);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN - tmem_shift);
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int col = col_base + c0;
        if (col >= p.n) break;  // warp-uniform
        float v[32];
        tmem_ld32(taddr0 + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        if (EPI == APH_EPI_QKV) {
          const int hidden = p.heads * 64;
          const int which = col / hidden;
          const int hc = col - which * hidden;
          const int h = hc >> 6;
          const int d0 = hc & 63;
          const float sc = which == 0 ? p.q_scale : 1.0f;
          {
            // bias (and, with a folded LayerNorm, the column sums) as 16-byte loads: col % 32 == 0 keeps them aligned.  As 32
            // scalar loads per chunk they were the largest single stall of this epilogue (profiles/r02_gemm_stalls.md).
            const float4* b4 = reinterpret_cast<const float4*>(vec_smem + c0);
            if (has_ln) {
              const float4* c4 = reinterpret_cast<const float4*>(vec_smem + 256 + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bv = b4[j];
                const float4 cv = c4[j];
                v[4 * j + 0] = fmaf(v[4 * j + 0], ln_rstd, fmaf(ln_nmr, cv.x, bv.x)) * sc;
                v[4 * j + 1] = fmaf(v[4 * j + 1], ln_rstd, fmaf(ln_nmr, cv.y, bv.y)) * sc;
                v[4 * j + 2] = fmaf(v[4 * j + 2], ln_rstd, fmaf(ln_nmr, cv.z, bv.z)) * sc;
                v[4 * j + 3] = fmaf(v[4 * j + 3], ln_rstd, fmaf(ln_nmr, cv.w, bv.w)) * sc;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bv = b4[j];
                v[4 * j + 0] = (v[4 * j + 0] + bv.x) * sc;
                v[4 * j + 1] = (v[4 * j + 1] + bv.y) * sc;
                v[4 * j + 2] = (v[4 * j + 2] + bv.z) * sc;
                v[4 * j + 3] = (v[4 * j + 3] + bv.w) * sc;
              }
            }
          }
          // Q / K / V [utterance*heads + head][frame][64]: the 64 columns of a head are two chunks; they go through the warp's
          // staging tile (32 frames x 128 bytes, 128B swizzle) and ONE TMA store — the map's frame axis ends at the utterance's
          // last frame, so rows that belong to the next utterance are clipped; the (few) lanes holding such rows copy their
          // 128 bytes out of the staging tile themselves (a TMA store may not start at a negative coordinate).  As 16-byte
          // row-strided stores from registers this epilogue was longer than the K = 1024 mainloop.
          {
            const int part = (c0 >> 5) & 1;
            if (part == 0 && store_pending) {
              if (lane == 0) bulk_store_wait_read<0>();
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(stage_row + (((4 * part + j) ^ sw) << 4)) = o;
            }
            if (part == 1) {
              fence_proxy_async_smem();
              __syncwarp();
              const long long grow0 = static_cast<long long>(mt) * kBM + quad * 32;  // batch == 1: first row of this warp
              const int utt0 = static_cast<int>(grow0 / p.len_period);
              const int tt0 = static_cast<int>(grow0 - static_cast<long long>(utt0) * p.len_period);
              if (lane == 0 && grow0 < p.a_rows) {
                const CUtensorMap* tm = which == 0 ? &tm_out : (which == 1 ? &tm_resid : &tm_copy);
                tma_store_3d(tm, stage_buf, 0, tt0, utt0 * p.heads + h);
              }
              store_pending = true;
              if (row_ok && utt != utt0) {  // this row opens the next utterance
                __nv_bfloat16* base = which == 0 ? p.q : (which == 1 ? p.kmat : p.vmat);
                uint4* d4 = reinterpret_cast<uint4*>(base + ((static_cast<long long>(utt) * p.heads + h) * p.len_period + tt) * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) d4[j] = *reinterpret_cast<const uint4*>(stage_row + ((j ^ sw) << 4));
              }
            }
          }
          if (which == 2 && p.vt != nullptr && row_ok) {  // optional transposed copy of V (nothing in the library reads it)
            const long long bh = static_cast<long long>(utt) * p.heads + h;
            __nv_bfloat16* dst = p.vt + (bh * 64 + d0) * p.t_v + tt;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[static_cast<long long>(j) * p.t_v] = __float2bfloat16(v[j]);
          }
        } else {
          const bool full_chunk = col + 32 <= p.n;
          if (has_ln) {
            const float4* b4 = reinterpret_cast<const float4*>(vec_smem + c0);
            const float4* c4 = reinterpret_cast<const float4*>(vec_smem + 256 + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = b4[j];
              const float4 cv = c4[j];
              v[4 * j + 0] = fmaf(v[4 * j + 0], ln_rstd, fmaf(ln_nmr, cv.x, bv.x));
              v[4 * j + 1] = fmaf(v[4 * j + 1], ln_rstd, fmaf(ln_nmr, cv.y, bv.y));
              v[4 * j + 2] = fmaf(v[4 * j + 2], ln_rstd, fmaf(ln_nmr, cv.z, bv.z));
              v[4 * j + 3] = fmaf(v[4 * j + 3], ln_rstd, fmaf(ln_nmr, cv.w, bv.w));
            }
          } else if (has_bias) {  // columns past n hold 0 in the staged vector
            const float4* b4 = reinterpret_cast<const float4*>(vec_smem + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = b4[j];
              v[4 * j + 0] = fmaf(v[4 * j + 0], out_scale, bv.x);
              v[4 * j + 1] = fmaf(v[4 * j + 1], out_scale, bv.y);
              v[4 * j + 2] = fmaf(v[4 * j + 2], out_scale, bv.z);
              v[4 * j + 3] = fmaf(v[4 * j + 3], out_scale, bv.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= out_scale;
          }
          if (has_aux && row_ok) {  // training forward: keep the pre-activation for the backward pass
            uint4* d4 = reinterpret_cast<uint4*>(p.aux_bf16 + grow * p.ld_aux + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (full_chunk || col + 8 * j + 8 <= p.n) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                d4[j] = o;
              }
            }
          }
          if (act == 1) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float2 x[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = make_float2(v[16 * g + 2 * j], v[16 * g + 2 * j + 1]);
              gelu_erf2_x8(x);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[16 * g + 2 * j] = x[j].x;
                v[16 * g + 2 * j + 1] = x[j].y;
              }
            }
          } else if (act == 2) {  // APH_ACT_RELU
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (act == 3) {  // APH_ACT_LEAKY_RELU (negative slope 0.01, nn.LeakyReLU's default)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
          }
          if (has_act_bwd && row_ok) {  // backward of GELU: dL/d(pre) = dL/d(act) * gelu'(pre)
            const uint4* s4 = reinterpret_cast<const uint4*>(p.gelu_bwd + grow * p.ld_gelu_bwd + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (full_chunk || col + 8 * j + 8 <= p.n) {
                const uint4 raw = s4[j];
                const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 pre = unpack_bf16x2(wds[e]);
                  if (p.act_bwd >= 2) {  // ReLU / LeakyReLU(0.01)
                    const float low = p.act_bwd == 3 ? 0.01f : 0.f;
                    v[8 * j + 2 * e + 0] *= pre.x > 0.f ? 1.f : low;
                    v[8 * j + 2 * e + 1] *= pre.y > 0.f ? 1.f : low;
                  } else {
                    v[8 * j + 2 * e + 0] *= gelu_erf_grad(pre.x);
                    v[8 * j + 2 * e + 1] *= gelu_erf_grad(pre.y);
                  }
                }
              }
            }
          }
          if (has_drop) {  // train-mode dropout of the branch output, before the residual joins
            const uint32_t key = drop_row_key(p.drop_seed, static_cast<uint32_t>(grow));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const uint32_t h = drop_hash(key, static_cast<uint32_t>(col >> 1) + j);
              v[2 * j + 0] = drop_keep(h, 0, p.drop_threshold) ? v[2 * j + 0] * p.drop_scale : 0.f;
              v[2 * j + 1] = drop_keep(h, 1, p.drop_threshold) ? v[2 * j + 1] * p.drop_scale : 0.f;
            }
          }
          if (resid_tma) {
            // this chunk's residual tile has landed (or lands now); copy the own row out, then hand the tile back to
            // the TMA engine for the next chunk.  Rows / columns outside the matrix arrive as zeros.
            mbar_wait(&resid_bar[warp], resid_phase);
            resid_phase ^= 1;
            float4 rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) rv[j] = *reinterpret_cast<const float4*>(resid_row + ((j ^ sw) << 4));
            fence_proxy_async_smem();
            __syncwarp();
            const int c0_next = c0 + 32;
            if (c0_next < c_end && col_base + c0_next < p.n && lane == 0) {
              mbar_arrive_expect_tx(&resid_bar[warp], 4096);
              tma_load_3d(resid_buf, &tm_resid, &resid_bar[warp], col_base + c0_next, t_tile_row0, b);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j + 0] += rv[j].x;
              v[4 * j + 1] += rv[j].y;
              v[4 * j + 2] += rv[j].z;
              v[4 * j + 3] += rv[j].w;
            }
          }
          if (row_ok) {
            if (plain_resid) {
              const float4* rs = reinterpret_cast<const float4*>(p.resid + grow * p.ld_resid + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (full_chunk || col + 4 * j + 4 <= p.n) {
                  const float4 rv = rs[j];
                  v[4 * j + 0] += rv.x;
                  v[4 * j + 1] += rv.y;
                  v[4 * j + 2] += rv.z;
                  v[4 * j + 3] += rv.w;
                }
              }
            }
            if (masked) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (direct_f32) {
              float4* d4 = reinterpret_cast<float4*>(p.out_f32 + grow * p.ld_f32 + col);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (full_chunk || col + 4 * j + 4 <= p.n)
                  d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            }
            if (direct_bf16) {
              uint4* d4 = reinterpret_cast<uint4*>(p.out_bf16 + grow * p.ld_bf16 + col);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (full_chunk || col + 8 * j + 8 <= p.n) {
                  uint4 o;
                  o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                  o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                  o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                  o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                  d4[j] = o;
                }
              }
            }
          }
          if (staged != 0 && mt < tiles_m) {  // warp-uniform
            if (masked || !row_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            const int t_row0 = (mt % p.m_tiles_per_batch) * kBM + quad * 32;
            if (staged == 1) {
              // fp32: this 32-column chunk is a full 128-byte row of the staging tile
              if (store_pending) {
                if (lane == 0) bulk_store_wait_read<0>();
                __syncwarp();
              }
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(stage_row + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              if (EPI == kEpiStoreResidStats) {
                // (the wait above covered the previous bf16-copy store as well: one bulk-group queue per thread)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  row_sum += v[j];
                  row_sq = fmaf(v[j], v[j], row_sq);
                }
                const int part = (c0 >> 5) & 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 o;
                  o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                  o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                  o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                  o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                  *reinterpret_cast<uint4*>(copy_row + (((4 * part + j) ^ sw) << 4)) = o;
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (p.split_k > 1) {
                  tma_reduce_add_3d(&tm_out, stage_buf, col, t_row0, b);
                } else {
                  tma_store_3d(&tm_out, stage_buf, col, t_row0, b);
                }
              }
              store_pending = true;
              if (EPI == kEpiStoreResidStats) {
                const int part = (c0 >> 5) & 1;
                const bool last_chunk = c0 + 32 >= c_end || col + 32 >= p.n;
                if ((part == 1 || last_chunk) && lane == 0) tma_store_3d(&tm_copy, copy_buf, col - 32 * part, t_row0, b);
              }
            } else {
              // bf16: two consecutive 32-column chunks fill the 128-byte rows (64 columns per store)
              const int part = (c0 >> 5) & 1;
              if (part == 0 && store_pending) {
                if (lane == 0) bulk_store_wait_read<0>();
                __syncwarp();
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4*>(stage_row + (((4 * part + j) ^ sw) << 4)) = o;
              }
              const bool last_chunk = c0 + 32 >= c_end || col + 32 >= p.n;
              if (part == 1 || last_chunk) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tma_store_3d(&tm_out, stage_buf, col - 32 * part, t_row0, b);
                store_pending = true;
              }
            }
          }
        }
      }
      if (EPI == kEpiStoreResidStats && row_ok && col_base + half * (BN / 2) < p.n)
        p.row_stats[grow * p.row_stats_slots + n_blk * 2 + half] = make_float2(row_sum, row_sq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the accumulator of BOTH CTAs must be drained before the leader overwrites it
        if (cta_rank == 0) {
          mbar_arrive(&tempty_bar[acc]);
        } else {
          mbar_arrive_remote(&tempty_bar[acc], 0);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (store_pending && lane == 0) bulk_store_wait_read<0>();  // smem must outlive the last TMA store's read
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's smem until it is done too
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

static bool taps_window_enabled() {
  static const bool enabled = [] {
    const char* e = getenv("APH_TAPS_WINDOW");
    return e == nullptr || atoi(e) != 0;
  }();
  return enabled;
}

template <int BN, int EPI>
static int launch_gemm(const aph_gemm_args* a, GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  CUtensorMap tm_a, tm_b;
  const int k_seq = a->k_seq;
  const int k_batch = a->k_batch > 0 ? a->k_batch : 1;
  if (a->a_mn_major) {
    // stored [segment][frame][output channel]: 64 x 64 boxes, frames past the segment read as zero
    const uint64_t dims[3] = {static_cast<uint64_t>(a->a_rows), static_cast<uint64_t>(k_seq), static_cast<uint64_t>(k_batch)};
    uint64_t seg_stride = static_cast<uint64_t>(a->a_batch_stride) * 2;
    if (k_batch == 1 || seg_stride == 0) seg_stride = static_cast<uint64_t>(a->a_row_stride) * 2 * static_cast<uint64_t>(k_seq);
    const uint64_t strides[2] = {static_cast<uint64_t>(a->a_row_stride) * 2, seg_stride};
    const uint32_t box[3] = {64, kBK, 1};
    int rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  } else {
    // a K-major A paired with an MN-major B contracts over k_seq columns only: clip the map so the tail reads zero
    // clip the map to the contraction length: columns past it (K-major B zero-filled there, or an MN-major B with
    // k_seq rows) read as zero instead of whatever follows in a wider A matrix
    const int k_len = a->b_mn_major ? k_seq : (a->mode == APH_GEMM_ROWS ? a->k : a->a_inner);
    const uint64_t inner = k_len < a->a_inner ? k_len : a->a_inner;
    const uint64_t dims[3] = {inner, static_cast<uint64_t>(a->a_rows), static_cast<uint64_t>(a->batch)};
    uint64_t batch_stride = static_cast<uint64_t>(a->a_batch_stride) * 2;
    if (a->batch == 1 && batch_stride == 0) batch_stride = static_cast<uint64_t>(a->a_row_stride) * 2;
    const uint64_t strides[2] = {static_cast<uint64_t>(a->a_row_stride) * 2, batch_stride};
    const uint32_t box[3] = {kBK, kBM, 1};
    int rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->a, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  if (a->b_mn_major) {
    const uint64_t row_stride = static_cast<uint64_t>(a->b_row_stride > 0 ? a->b_row_stride : a->n) * 2;
    const uint64_t dims[3] = {static_cast<uint64_t>(a->n), static_cast<uint64_t>(k_seq), static_cast<uint64_t>(k_batch)};
    uint64_t seg_stride = static_cast<uint64_t>(a->b_seg_stride) * 2;
    if (k_batch == 1 || seg_stride == 0) seg_stride = row_stride * static_cast<uint64_t>(k_seq);
    const uint64_t strides[2] = {row_stride, seg_stride};
    const uint32_t box[3] = {64, kBK, 1};
    int rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  } else {
    const uint64_t row_stride = static_cast<uint64_t>(a->b_row_stride > 0 ? a->b_row_stride : a->k) * 2;
    const uint64_t dims[2] = {static_cast<uint64_t>(a->k), static_cast<uint64_t>(a->n)};
    const uint64_t strides[1] = {row_stride};
    const uint32_t box[2] = {kBK, static_cast<uint32_t>(BN / 2)};  // each CTA of a cluster fetches half the B tile
    int rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  CUtensorMap tm_out;
  memset(&tm_out, 0, sizeof(tm_out));
  if (p.staged != 0) {
    const bool f32 = p.staged == 1;
    const uint64_t es = f32 ? 4 : 2;
    const uint64_t ld = static_cast<uint64_t>(f32 ? a->ld_f32 : a->ld_bf16);
    const uint64_t out_batches = p.diag_taps > 0 ? static_cast<uint64_t>(p.diag_taps) : static_cast<uint64_t>(a->batch);
    const uint64_t dims[3] = {static_cast<uint64_t>(p.n), static_cast<uint64_t>(a->a_rows), out_batches};
    uint64_t batch_stride = static_cast<uint64_t>(a->out_batch_rows) * ld * es;
    if (out_batches == 1 || batch_stride == 0) batch_stride = static_cast<uint64_t>(a->a_rows) * ld * es;
    const uint64_t strides[2] = {ld * es, batch_stride};
    const uint32_t box[3] = {f32 ? 32u : 64u, 32u, 1u};
    int rc = encode_tmap(&tm_out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                         f32 ? static_cast<const void*>(a->out_f32) : a->out_bf16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  CUtensorMap tm_resid;
  memset(&tm_resid, 0, sizeof(tm_resid));
  if (EPI == kEpiStoreResidTma || EPI == kEpiStoreResidStats) {
    const uint64_t ld = static_cast<uint64_t>(a->ld_resid);
    const uint64_t dims[3] = {static_cast<uint64_t>(p.n), static_cast<uint64_t>(a->a_rows), static_cast<uint64_t>(a->batch)};
    uint64_t batch_stride = static_cast<uint64_t>(a->out_batch_rows) * ld * 4;
    if (a->batch == 1 || batch_stride == 0) batch_stride = static_cast<uint64_t>(a->a_rows) * ld * 4;
    const uint64_t strides[2] = {ld * 4, batch_stride};
    const uint32_t box[3] = {32u, 32u, 1u};
    int rc = encode_tmap(&tm_resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a->resid, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  CUtensorMap tm_copy;
  memset(&tm_copy, 0, sizeof(tm_copy));
  if (EPI == APH_EPI_QKV) {  // Q, K, V [utterance*heads + head][frame][64] ride in the three output-side map slots
    const uint64_t n_utt = static_cast<uint64_t>(ceil_div(a->a_rows, a->len_period));
    const uint64_t dims[3] = {64, static_cast<uint64_t>(a->len_period), n_utt * static_cast<uint64_t>(a->heads)};
    const uint64_t strides[2] = {128, static_cast<uint64_t>(a->len_period) * 128};
    const uint32_t box[3] = {64u, 32u, 1u};
    int rc = encode_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->q, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc == APH_OK) rc = encode_tmap(&tm_resid, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->kmat, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc == APH_OK) rc = encode_tmap(&tm_copy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->vmat, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != APH_OK) return rc;
  }
  if (EPI == kEpiStoreResidStats) {  // bf16 copy of the fp32 output: 64-column x 32-row boxes
    const uint64_t ld = static_cast<uint64_t>(a->ld_bf16);
    const uint64_t dims[3] = {static_cast<uint64_t>(p.n), static_cast<uint64_t>(a->a_rows), static_cast<uint64_t>(a->batch)};
    uint64_t batch_stride = static_cast<uint64_t>(a->out_batch_rows) * ld * 2;
    if (a->batch == 1 || batch_stride == 0) batch_stride = static_cast<uint64_t>(a->a_rows) * ld * 2;
    const uint64_t strides[2] = {ld * 2, batch_stride};
    const uint32_t box[3] = {BN >= 128 ? 64u : 32u, 32u, 1u};
    int rc = encode_tmap(&tm_copy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->out_bf16, dims, strides, box,
                         BN >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc != APH_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  p.idesc = umma_idesc_bf16(2 * kBM, BN) | (p.a_mn ? kIdescAMnMajor : 0u) | (p.b_mn ? kIdescBMnMajor : 0u);
  const int m_pairs = (p.m_tiles_per_batch * p.batch + 1) / 2;
  const int max_clusters = sm_count() / 2;
  // split K when the output has too few tiles to fill the machine (weight gradients: K = frames is the long axis)
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  if (a->a_mn_major && p.diag_taps == 0 && p.staged == 1 && a->bias == nullptr && a->resid == nullptr && !a->gelu &&
      a->aux_bf16 == nullptr && a->gelu_bwd == nullptr && a->out_bf16 == nullptr) {
    const int tiles = m_pairs * p.n_tiles;
    int want = max_clusters / tiles;                  // splits that still fit one wave
    const int max_by_k = p.k_blocks / 8;              // keep >= 8 k-blocks (512 frames) per split
    if (want > max_by_k) want = max_by_k;
    if (want > 1) {
      p.kb_per_split = ceil_div(p.k_blocks, want);
      p.split_k = ceil_div(p.k_blocks, p.kb_per_split);
      // partial tiles are reduce-added: start from zero
      if (a->ld_f32 == a->n) {
        APH_CUDA_CHECK(cudaMemsetAsync(a->out_f32, 0, sizeof(float) * static_cast<size_t>(a->a_rows) * a->n, stream));
      } else {
        APH_CUDA_CHECK(cudaMemset2DAsync(a->out_f32, sizeof(float) * a->ld_f32, 0, sizeof(float) * a->n, a->a_rows, stream));
      }
    }
  }
  int total_work = p.diag_taps > 0 ? m_pairs * p.diag_taps : m_pairs * p.n_tiles * p.split_k;
  p.tail_first = total_work;
  // tail split (decode_work): the R tiles of a partly filled last wave are computed as 2 R half-width items
  CUtensorMap tm_b_narrow;
  memset(&tm_b_narrow, 0, sizeof(tm_b_narrow));
  if (BN == 256 && (EPI == APH_EPI_STORE || EPI == kEpiStoreResidTma || EPI == APH_EPI_QKV) && p.split_k == 1 && p.diag_taps == 0 && !p.a_mn &&
      p.mode == APH_GEMM_ROWS && p.n % 128 == 0 && tail_split_enabled()) {
    const int full = (total_work / max_clusters) * max_clusters;
    const int rest = total_work - full;
    if (rest > 0 && 2 * rest <= max_clusters) {
      if (!a->b_mn_major) {  // a K-major B is fetched as [BN/4 rows][64] boxes by the half-width items
        const uint64_t row_stride = static_cast<uint64_t>(a->b_row_stride > 0 ? a->b_row_stride : a->k) * 2;
        const uint64_t dims[2] = {static_cast<uint64_t>(a->k), static_cast<uint64_t>(a->n)};
        const uint64_t strides[1] = {row_stride};
        const uint32_t box[2] = {kBK, static_cast<uint32_t>(BN / 4)};
        int rc = encode_tmap(&tm_b_narrow, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != APH_OK) return rc;
      }
      p.idesc_narrow = umma_idesc_bf16(2 * kBM, BN / 2) | (p.a_mn ? kIdescAMnMajor : 0u) | (p.b_mn ? kIdescBMnMajor : 0u);
      p.tail_first = full;
      total_work = full + 2 * rest;
    }
  }
  // sliding A window for the tap GEMM over 64-channel groups (see the producer)
  CUtensorMap tm_a_win;
  memset(&tm_a_win, 0, sizeof(tm_a_win));
  p.taps_window = 0;
  p.win_bytes = 0;
  if (BN == 64 && a->mode == APH_GEMM_TAPS && p.taps_span == 64 && !p.a_mn && !p.b_mn && p.split_k == 1 && taps_window_enabled()) {
    const int window = p.k_blocks % 64 == 0 ? 64 : (p.k_blocks % 32 == 0 ? 32 : 0);
    if (window > 0 && kWinStages <= Cfg::kStages && window % kTapsPerStage == 0 &&
        2 * (window + 128) * 128 + kWinStages * kTapsPerStage * Cfg::kBBytes <= Cfg::kStages * Cfg::kStageBytes) {
      const uint64_t inner = static_cast<uint64_t>(a->a_inner);
      const uint64_t dims[3] = {inner, static_cast<uint64_t>(a->a_rows), static_cast<uint64_t>(a->batch)};
      uint64_t batch_stride = static_cast<uint64_t>(a->a_batch_stride) * 2;
      if (a->batch == 1 && batch_stride == 0) batch_stride = static_cast<uint64_t>(a->a_row_stride) * 2;
      const uint64_t strides[2] = {static_cast<uint64_t>(a->a_row_stride) * 2, batch_stride};
      const uint32_t box[3] = {kBK, static_cast<uint32_t>(window + 128), 1};
      int rc = encode_tmap(&tm_a_win, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->a, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != APH_OK) return rc;
      p.taps_window = window;
      p.win_bytes = (window + 128) * 128;
    }
  }
  p.total_work = total_work;
  const int grid = 2 * (total_work < max_clusters ? total_work : max_clusters);
  APH_CUDA_CHECK(launch_pdl(gemm_bf16_kernel<BN, EPI>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tm_a, tm_b, tm_out, tm_resid, tm_copy, tm_b_narrow, tm_a_win, p));
  APH_POST_LAUNCH(1);
  return APH_OK;
}

}  // namespace aph

extern "C" int aph_gemm_bf16(const aph_gemm_args* a, void* stream_) {
  using namespace aph;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(a != nullptr, "null args");
  APH_REQUIRE(a->a != nullptr && a->b != nullptr, "null operand");
  const bool mn = a->a_mn_major || a->b_mn_major;
  APH_REQUIRE(mn || (a->k > 0 && a->k % 8 == 0), "k must be a positive multiple of 8 (the tail of the last 64-wide block is TMA zero fill)");
  APH_REQUIRE(a->n > 0 && a->n % 8 == 0, "n must be a positive multiple of 8");
  APH_REQUIRE(a->a_rows > 0 && a->batch > 0, "empty A");
  APH_REQUIRE(a->a_row_stride % 8 == 0 && a->a_batch_stride % 8 == 0, "A strides must be multiples of 8 elements");
  APH_REQUIRE(a->b_row_stride % 8 == 0 && a->b_seg_stride % 8 == 0, "B strides must be multiples of 8 elements");
  APH_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->b) & 15) == 0,
              "operands must be 16-byte aligned");
  APH_REQUIRE(a->mode == APH_GEMM_ROWS || a->mode == APH_GEMM_TAPS || a->mode == APH_GEMM_DIAG_TAPS, "bad mode");
  APH_REQUIRE(!a->a_mn_major || a->b_mn_major, "an MN-major A needs an MN-major B (weight-gradient form)");
  APH_REQUIRE(!mn || (a->k_seq > 0 && a->k_batch >= 0), "MN-major operands need k_seq");
  APH_REQUIRE(!mn || a->mode != APH_GEMM_TAPS, "taps mode takes K-major operands");
  APH_REQUIRE(!a->a_mn_major || (a->batch == 1 && a->a_rows % 8 == 0), "MN-major A: batch 1, a_rows % 8 == 0");
  APH_REQUIRE(!(a->b_mn_major && !a->a_mn_major) || a->k_batch <= 1, "K-major A with MN-major B: one segment");
  APH_REQUIRE(!a->aux_bf16 || (a->ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux_bf16) & 15) == 0),
              "aux output must be 16-byte aligned with ld % 8 == 0");
  APH_REQUIRE(!a->gelu_bwd || (a->ld_gelu_bwd % 8 == 0 && (reinterpret_cast<uintptr_t>(a->gelu_bwd) & 15) == 0),
              "gelu_bwd source must be 16-byte aligned with ld % 8 == 0");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles_per_batch = ceil_div(a->a_rows, kBM);
  p.batch = a->batch;
  if (mn) {
    p.k_seq_blocks = ceil_div(a->k_seq, kBK);
    p.k_blocks = p.k_seq_blocks * (a->k_batch > 0 ? a->k_batch : 1);
  } else {
    p.k_seq_blocks = 1;
    p.k_blocks = ceil_div(a->k, kBK);
  }
  p.a_mn = a->a_mn_major ? 1 : 0;
  p.b_mn = a->b_mn_major ? 1 : 0;
  p.b_k_shift = a->b_k_shift;
  p.a_rows = a->a_rows;
  p.mode = a->mode;
  p.tap_pad = a->tap_pad;
  p.n = a->n;
  p.gelu = a->gelu;
  APH_REQUIRE(a->gelu >= 0 && a->gelu <= 3, "gelu: 0 none, 1 GELU(erf), 2 ReLU, 3 LeakyReLU(0.01)");
  p.scale = a->scale;
  p.bias = a->bias;
  p.resid = a->resid;
  p.ld_resid = a->ld_resid;
  p.out_f32 = a->out_f32;
  p.ld_f32 = a->ld_f32;
  p.out_bf16 = static_cast<__nv_bfloat16*>(a->out_bf16);
  p.ld_bf16 = a->ld_bf16;
  p.out_batch_rows = a->out_batch_rows;
  p.lengths = a->lengths;
  p.len_period = a->len_period;
  p.q = static_cast<__nv_bfloat16*>(a->q);
  p.kmat = static_cast<__nv_bfloat16*>(a->kmat);
  p.vt = static_cast<__nv_bfloat16*>(a->vt);
  p.vmat = static_cast<__nv_bfloat16*>(a->vmat);
  p.drop_threshold = a->drop_threshold;
  p.drop_seed = a->drop_seed;
  p.drop_scale = a->drop_scale;
  p.act_bwd = a->act_bwd;
  APH_REQUIRE(a->act_bwd >= 0 && a->act_bwd <= 3, "act_bwd: 0/1 GELU, 2 ReLU, 3 LeakyReLU");
  APH_REQUIRE(a->drop_threshold < 65536u, "drop_threshold is 16 bits (p < 1)");
  APH_REQUIRE(a->drop_threshold == 0 || (a->epilogue == APH_EPI_STORE && !a->a_mn_major), "dropout: store epilogue of a forward GEMM only");
  p.heads = a->heads;
  p.t_v = a->t_v;
  p.q_scale = a->q_scale;
  p.aux_bf16 = static_cast<__nv_bfloat16*>(a->aux_bf16);
  p.ld_aux = a->ld_aux;
  p.gelu_bwd = static_cast<const __nv_bfloat16*>(a->gelu_bwd);
  p.ld_gelu_bwd = a->ld_gelu_bwd;
  p.staged = 0;
  p.split_k = 1;
  p.kb_per_split = p.k_blocks;
  // LayerNorm folded into the GEMMs around it
  p.row_stats = reinterpret_cast<float2*>(a->row_stats);
  p.row_stats_slots = a->row_stats_slots;
  p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
  p.ln_slots = a->ln_slots;
  p.ln_colsum = a->ln_colsum;
  p.ln_inv_cols = a->ln_cols > 0 ? 1.0f / static_cast<float>(a->ln_cols) : 0.f;
  p.ln_eps = a->ln_eps;
  const bool stats_out = a->row_stats != nullptr;
  const bool stats_in = a->ln_stats != nullptr;
  APH_REQUIRE(!stats_in || (a->ln_slots > 0 && a->ln_slots <= 16 && a->ln_slots % 2 == 0 && a->ln_cols > 0 && a->ln_colsum && a->bias && !mn && a->mode == APH_GEMM_ROWS && a->batch == 1 &&
                            a->n % 32 == 0 && a->scale == 1.0f && (reinterpret_cast<uintptr_t>(a->ln_colsum) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(a->ln_stats) & 15) == 0),
              "ln_stats: plain K-major GEMM with bias, scale 1, n % 32 == 0, batch 1, even ln_slots <= 16, ln_cols / ln_colsum set");
  APH_REQUIRE(!stats_out || (a->resid && a->out_f32 && a->out_bf16 && !mn && a->mode == APH_GEMM_ROWS && a->epilogue == APH_EPI_STORE &&
                             a->n > 128 && a->n % 64 == 0 && a->row_stats_slots == 2 * ceil_div(a->n, 256) &&
                             (reinterpret_cast<uintptr_t>(a->row_stats) & 7) == 0),
              "row_stats: fp32 output with residual and a bf16 copy, n a multiple of 64 above 128, row_stats_slots == 2 * ceil(n / 256)");

  if (a->mode == APH_GEMM_DIAG_TAPS) {
    // dW of the grouped positional conv: for tap j and 256-channel block q,
    //   out[j][q*256 + m][c] = sum_frames A[frame][q*256 + m] * B[frame + j - tap_pad][q*256 + c]
    APH_REQUIRE(a->a_mn_major && a->b_mn_major, "diag-taps mode takes MN-major operands");
    APH_REQUIRE(a->epilogue == APH_EPI_STORE && a->out_f32 && !a->out_bf16 && !a->bias && !a->resid, "diag-taps: fp32 output only");
    APH_REQUIRE(a->n_taps > 0 && a->a_rows % 256 == 0 && a->n == a->a_rows, "diag-taps: square, channels % 256 == 0");
    APH_REQUIRE(a->ld_f32 == 256 && a->out_batch_rows == a->a_rows, "diag-taps: output [n_taps][channels][256]");
    p.diag_taps = a->n_taps;
    p.n_tiles = a->a_rows / 256;
    p.n = 256;
    p.staged = 1;
    return launch_gemm<256, APH_EPI_STORE>(a, p, stream);
  }
  if (a->mode == APH_GEMM_TAPS) {
    // one 64-channel group per N tile; k-block j is tap j of that group
    APH_REQUIRE(a->n % 64 == 0 && a->a_inner == a->n, "taps mode: n == channels, multiple of 64");
    p.taps_span = a->taps_span > 0 ? a->taps_span : 64;
    APH_REQUIRE(p.taps_span % 64 == 0 && a->n % p.taps_span == 0 && a->k % p.taps_span == 0, "taps mode: taps_span a multiple of 64 dividing n and k");
    APH_REQUIRE(a->epilogue == APH_EPI_STORE, "taps mode supports the store epilogue only");
    p.n_tiles = a->n / 64;
    p.staged = a->out_f32 ? 1 : 0;  // 64-column tiles: only the fp32 staging granularity (32 columns) fits
    if (a->resid && p.staged == 1) return launch_gemm<64, kEpiStoreResidTma>(a, p, stream);
    return launch_gemm<64, APH_EPI_STORE>(a, p, stream);
  }
  APH_REQUIRE(mn || a->a_inner >= a->k, "A rows shorter than k");

  if (a->epilogue == APH_EPI_QKV) {
    APH_REQUIRE(!mn, "qkv epilogue takes K-major operands");
    APH_REQUIRE(a->q && a->kmat && a->vmat && a->bias, "qkv epilogue needs q/k/v/bias");
    APH_REQUIRE(a->batch == 1 && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "qkv epilogue: batch 1, bias 16-byte aligned");
    APH_REQUIRE(((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->kmat) | reinterpret_cast<uintptr_t>(a->vmat)) & 15) == 0,
                "qkv epilogue: q / k / v must be 16-byte aligned");
    APH_REQUIRE(a->heads > 0 && a->n == 3 * a->heads * 64, "qkv epilogue: n == 3*heads*64");
    APH_REQUIRE(a->len_period > 0 && (!a->vt || (a->t_v % 8 == 0 && a->t_v >= a->len_period)), "qkv epilogue: bad lengths");
    APH_REQUIRE((a->heads * 64) % 256 == 0, "qkv epilogue: hidden must be a multiple of 256");
    p.n_tiles = ceil_div(a->n, 256);
    return launch_gemm<256, APH_EPI_QKV>(a, p, stream);
  }
  APH_REQUIRE(a->epilogue == APH_EPI_STORE, "bad epilogue");
  APH_REQUIRE(a->out_f32 || a->out_bf16, "no output");
  APH_REQUIRE(!a->out_f32 || (a->ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out_f32) & 15) == 0),
              "fp32 output must be 16-byte aligned with ld % 4 == 0");
  APH_REQUIRE(!a->out_bf16 || (a->ld_bf16 % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out_bf16) & 15) == 0),
              "bf16 output must be 16-byte aligned with ld % 8 == 0");
  APH_REQUIRE(!a->resid || (a->ld_resid % 4 == 0 && (reinterpret_cast<uintptr_t>(a->resid) & 15) == 0),
              "residual must be 16-byte aligned with ld % 4 == 0");
  APH_REQUIRE(!a->lengths || a->len_period > 0, "lengths need len_period");
  // primary output goes through the staged TMA store (fp32 wins when both are requested)
  p.staged = a->out_f32 ? 1 : 2;
  if (a->n > 128) {
    // Wave quantisation (80 tiles of 256 x 256 on 74 cluster slots for the 4.9 k x 1024 GEMMs of a training step) is NOT
    // cured by 256 x 128 tiles: measured on the training step, they cost ~0.7 of a full tile each and the step gets 4 %
    // slower (profiles/r01_gemm_bn128_ab.md).  Kept behind APH_GEMM_BN128=1 for experiments.
    static const bool narrow_requested = [] {
      const char* e = getenv("APH_GEMM_BN128");
      return e != nullptr && atoi(e) == 1;
    }();
    const bool narrow = narrow_requested && !a->a_mn_major;
    p.n_tiles = ceil_div(a->n, narrow ? 128 : 256);
    // fp32 output with a residual (out-proj, FFN2, gradient accumulation): the residual comes in through TMA
    if (stats_out) {
      p.n_tiles = ceil_div(a->n, 256);
      return launch_gemm<256, kEpiStoreResidStats>(a, p, stream);
    }
    if (a->resid && p.staged == 1 && !a->a_mn_major)
      return narrow ? launch_gemm<128, kEpiStoreResidTma>(a, p, stream) : launch_gemm<256, kEpiStoreResidTma>(a, p, stream);
    return narrow ? launch_gemm<128, APH_EPI_STORE>(a, p, stream) : launch_gemm<256, APH_EPI_STORE>(a, p, stream);
  } else if (a->n > 64 || a->b_mn_major) {  // an MN-major B tile needs at least one 64-column chunk per CTA
    p.n_tiles = 1;
    return launch_gemm<128, APH_EPI_STORE>(a, p, stream);
  }
  p.n_tiles = 1;
  if (p.staged == 2) p.staged = 0;  // 64-column tiles: bf16 staging needs 64 columns per warp
  return launch_gemm<64, APH_EPI_STORE>(a, p, stream);
}
