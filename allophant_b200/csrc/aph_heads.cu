// allophant_b200 — classifier-head kernels (all HBM-/latency-bound, no tensor cores).
//
//   aph_compose_embeddings    EmbeddingCompositionLayer gather-sum     acoustic_model.py:219-232
//   aph_log_softmax_heads     log_softmax of every head in one launch  acoustic_model.py:1051-1052, loss_functions.py:27
//                             (+ per-frame argmax / max log-prob for greedy decoding, predictions.py:195)
//   aph_log_softmax_wide      same for one wide head (large inventories), warp per frame
//   aph_dependency_softmax    softmax of dependency logits -> bf16 classifier input   acoustic_model.py:497-514
//   aph_argmax_rows           torch.max(log_emissions, -1)             predictions.py:195
//   aph_ctc_greedy_collapse   unique_consecutive / blank removal / timesteps / score   predictions.py:196-206
#include "aph_common.cuh"

namespace aph {

// ---------------------------------------------------------------------------
// composed[0]   = W[0]                                  (blank embedding)
// composed[1+v] = sum_f W[tfi[v][f] + offsets[f]]       (EmbeddingBag mode="sum", fp32, f ascending)
// rows [1+V, rows_out) are zero (GEMM padding).  Output bf16 [rows_out][E] (+ optional fp32 copy).
// ---------------------------------------------------------------------------
__global__ void compose_embeddings_kernel(const float* __restrict__ weight, int n_categories, int E,
                                          const long long* __restrict__ tfi, const long long* __restrict__ offsets, int V,
                                          int F, int rows_out, __nv_bfloat16* __restrict__ out_bf16,
                                          float* __restrict__ out_f32, int* __restrict__ err_flag) {
  const int row = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float acc = 0.f;
    if (row == 0) {
      acc = weight[e];
    } else if (row <= V) {
      const long long* idx = tfi + static_cast<long long>(row - 1) * F;
      for (int f = 0; f < F; ++f) {
        const long long cat = idx[f] + (offsets ? offsets[f] : 0);
        if (cat < 0 || cat >= n_categories) {
          if (e == 0) atomicExch(err_flag, 1);
          continue;
        }
        acc += weight[cat * E + e];
      }
    }
    if (out_bf16) out_bf16[static_cast<long long>(row) * E + e] = __float2bfloat16(acc);
    if (out_f32) out_f32[static_cast<long long>(row) * E + e] = acc;
  }
}

// ---------------------------------------------------------------------------
// Multi-head log-softmax for narrow heads.  logits: fp32 [rows][ld]; head h
// occupies columns [col_off[h], col_off[h] + width[h]).  Output of head h is a
// contiguous fp32 [rows][width[h]] block at out + out_off[h].
// A block stages kRowsPerBlock rows of the packed logits in shared memory
// (coalesced 16-byte loads), then each thread handles (row, head) pairs with
// rows fastest so the per-head stores are contiguous across the warp.
// ---------------------------------------------------------------------------
constexpr int kLsmRows = 32;

__global__ void __launch_bounds__(256) log_softmax_heads_kernel(const float* __restrict__ logits, long long ld, long long rows,
                                                                int col_lo, int col_span, const int* __restrict__ col_off,
                                                                const int* __restrict__ width,
                                                                const long long* __restrict__ out_off, int n_heads,
                                                                float* __restrict__ out, int* __restrict__ argmax_out,
                                                                float* __restrict__ maxlp_out) {
  extern __shared__ float tile[];  // [kLsmRows][col_span | 1]
  const int stride = col_span | 1;
  const long long row0 = static_cast<long long>(blockIdx.x) * kLsmRows;
  const int n_rows = static_cast<int>(min(static_cast<long long>(kLsmRows), rows - row0));
  // col_lo and col_span are multiples of 4 and ld % 4 == 0: float4 loads are aligned
  const int vec_per_row = col_span >> 2;
  for (int i = threadIdx.x; i < n_rows * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c4 = i - r * vec_per_row;
    const float4 v = *reinterpret_cast<const float4*>(logits + (row0 + r) * ld + col_lo + 4 * c4);
    float* d = tile + r * stride + 4 * c4;
    d[0] = v.x;
    d[1] = v.y;
    d[2] = v.z;
    d[3] = v.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_heads * kLsmRows; i += blockDim.x) {
    const int h = i / kLsmRows, r = i - h * kLsmRows;
    if (r >= n_rows) continue;
    const int w = width[h];
    const float* src = tile + r * stride + (col_off[h] - col_lo);
    float m = src[0];
    int am = 0;
    for (int c = 1; c < w; ++c) {
      const float v = src[c];
      if (v > m) {
        m = v;
        am = c;
      }
    }
    float s = 0.f;
    for (int c = 0; c < w; ++c) s += expf(src[c] - m);
    const float ls = logf(s);
    float* dst = out + out_off[h] + (row0 + r) * w;
    for (int c = 0; c < w; ++c) dst[c] = (src[c] - m) - ls;  // same operation order as ATen's log_softmax
    if (argmax_out) argmax_out[static_cast<long long>(h) * rows + row0 + r] = am;
    if (maxlp_out) maxlp_out[static_cast<long long>(h) * rows + row0 + r] = -ls;
  }
}

// ---------------------------------------------------------------------------
// Wide head (e.g. composed phoneme logits over a full inventory): one warp per
// frame, the row is staged in shared memory so HBM sees exactly one read and
// one write per element.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) log_softmax_wide_kernel(const float* __restrict__ logits, long long ld, long long rows,
                                                               int width, float* __restrict__ out, long long ld_out,
                                                               int* __restrict__ argmax_out, float* __restrict__ maxlp_out,
                                                               int warps_per_block) {
  extern __shared__ float cache[];  // [warps_per_block][width]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps_per_block) return;
  float* buf = cache + static_cast<long long>(warp) * width;
  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_block + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * warps_per_block) {
    const float* src = logits + row * ld;
    float m = -INFINITY;
    int am = 0;
    for (int c = lane; c < width; c += 32) {
      const float v = __ldg(src + c);
      buf[c] = v;
      if (v > m) {
        m = v;
        am = c;
      }
    }
    // warp argmax with lowest-index tie-break (torch.max semantics)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o);
      const int ao = __shfl_xor_sync(0xffffffffu, am, o);
      if (mo > m || (mo == m && ao < am)) {
        m = mo;
        am = ao;
      }
    }
    float s = 0.f;
    for (int c = lane; c < width; c += 32) s += expf(buf[c] - m);
    s = warp_sum(s);
    const float ls = logf(s);
    float* dst = out + row * ld_out;
    for (int c = lane; c < width; c += 32) dst[c] = (buf[c] - m) - ls;
    if (lane == 0) {
      if (argmax_out) argmax_out[row] = am;
      if (maxlp_out) maxlp_out[row] = -ls;
    }
    __syncwarp();
  }
}

// Same contract, rows moved by the TMA engine: every warp double-buffers whole rows in shared memory
// (cp.async.bulk global->smem signalled on an mbarrier), makes one online max/sum pass and one
// in-place normalisation pass over smem with 16-byte accesses, and writes the row back with a bulk
// smem->global store.  HBM sees each element once in and once out; no register staging, the loads
// of row i+1 and the store of row i-1 overlap the arithmetic of row i.  Needs 16-byte aligned rows.
constexpr int kLsmBulkWarps = 4;

__global__ void __launch_bounds__(kLsmBulkWarps * 32) log_softmax_wide_bulk_kernel(const float* __restrict__ logits, long long ld,
                                                                                   long long rows, int width,
                                                                                   float* __restrict__ out, long long ld_out,
                                                                                   int* __restrict__ argmax_out,
                                                                                   float* __restrict__ maxlp_out) {
  extern __shared__ __align__(16) uint8_t lsm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t row_bytes = static_cast<uint32_t>(width) * 4u;
  const uint32_t buf_bytes = (row_bytes + 127u) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lsm_smem);  // [warps][2]
  float* bufs = reinterpret_cast<float*>(lsm_smem + 128);
  float* buf0 = bufs + static_cast<size_t>(warp) * 2 * (buf_bytes / 4);
  uint64_t* bar = bars + warp * 2;
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  const long long first = static_cast<long long>(blockIdx.x) * kLsmBulkWarps + warp;
  const long long step = static_cast<long long>(gridDim.x) * kLsmBulkWarps;
  if (first >= rows) return;
  if (lane == 0) {
    mbar_arrive_expect_tx(&bar[0], row_bytes);
    bulk_load_1d(buf0, logits + first * ld, row_bytes, &bar[0]);
  }
  const int n_vec = width >> 2;
  int it = 0;
  for (long long row = first; row < rows; row += step, ++it) {
    const int cur = it & 1;
    float* buf = buf0 + cur * (buf_bytes / 4);
    const long long next = row + step;
    if (next < rows && lane == 0) {
      bulk_store_wait_read<0>();  // the store that last read the other buffer has drained it
      mbar_arrive_expect_tx(&bar[cur ^ 1], row_bytes);
      bulk_load_1d(buf0 + (cur ^ 1) * (buf_bytes / 4), logits + next * ld, row_bytes, &bar[cur ^ 1]);
    }
    mbar_wait(&bar[cur], (it >> 1) & 1);
    // pass 1: online max / sum (4 independent chains per lane), argmax with lowest-index tie-break
    float m = -INFINITY, s = 0.f;
    int am = 0;
    const float4* v4 = reinterpret_cast<const float4*>(buf);
    for (int i = lane; i < n_vec; i += 32) {
      const float4 v = v4[i];
      const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (m4 > m) {
        s *= __expf(m - m4);
        m = m4;
        am = 4 * i + (v.x == m4 ? 0 : (v.y == m4 ? 1 : (v.z == m4 ? 2 : 3)));
      }
      s += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o);
      const float so = __shfl_xor_sync(0xffffffffu, s, o);
      const int ao = __shfl_xor_sync(0xffffffffu, am, o);
      const float mn = fmaxf(m, mo);
      s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (mo == -INFINITY ? 0.f : so * __expf(mo - mn));
      if (mo > m || (mo == m && ao < am)) am = ao;
      m = mn;
    }
    const float ls = logf(s);
    // pass 2: normalise in place
    float4* w4 = reinterpret_cast<float4*>(buf);
    for (int i = lane; i < n_vec; i += 32) {
      float4 v = w4[i];
      v.x = (v.x - m) - ls;
      v.y = (v.y - m) - ls;
      v.z = (v.z - m) - ls;
      v.w = (v.w - m) - ls;
      w4[i] = v;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_store_1d(out + row * ld_out, buf, row_bytes);
      if (argmax_out) argmax_out[row] = am;
      if (maxlp_out) maxlp_out[row] = -ls;
    }
  }
  if (lane == 0) bulk_store_wait_read<0>();
}

// ---------------------------------------------------------------------------
// softmax(dependency logits[..., skip:]) -> bf16 columns of the next classifier's input
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dependency_softmax_kernel(const float* __restrict__ logits, long long ld, long long rows,
                                                                 const int* __restrict__ col_off, const int* __restrict__ width,
                                                                 const int* __restrict__ dst_col, int n_deps, int skip,
                                                                 __nv_bfloat16* __restrict__ dst, long long ld_dst) {
  const long long total = rows * n_deps;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / n_deps;
    const int d = static_cast<int>(i - row * n_deps);
    const float* src = logits + row * ld + col_off[d] + skip;
    const int w = width[d] - skip;
    float m = -INFINITY;
    for (int c = 0; c < w; ++c) m = fmaxf(m, src[c]);
    float s = 0.f;
    for (int c = 0; c < w; ++c) s += expf(src[c] - m);
    const float inv = 1.0f / s;
    __nv_bfloat16* o = dst + row * ld_dst + dst_col[d];
    for (int c = 0; c < w; ++c) o[c] = __float2bfloat16(expf(src[c] - m) * inv);
  }
}

// ---------------------------------------------------------------------------
// argmax over the class axis of [rows][width] (arbitrary row stride), lowest index wins ties
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, long long ld, long long rows, int width,
                                                          int* __restrict__ argmax_out, float* __restrict__ max_out) {
  const int lane = threadIdx.x & 31;
  if (width <= 32) {
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (row >= rows) return;
    const float* src = x + row * ld;
    float m = src[0];
    int am = 0;
    for (int c = 1; c < width; ++c) {
      const float v = src[c];
      if (v > m) {
        m = v;
        am = c;
      }
    }
    argmax_out[row] = am;
    max_out[row] = m;
  } else {
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* src = x + row * ld;
    float m = -INFINITY;
    int am = 0;
    for (int c = lane; c < width; c += 32) {
      const float v = __ldg(src + c);
      if (v > m) {
        m = v;
        am = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o);
      const int ao = __shfl_xor_sync(0xffffffffu, am, o);
      if (mo > m || (mo == m && ao < am)) {
        m = mo;
        am = ao;
      }
    }
    if (lane == 0) {
      argmax_out[row] = am;
      max_out[row] = m;
    }
  }
}

// ---------------------------------------------------------------------------
// Greedy CTC collapse: one warp per (head, utterance) sequence of argmax ids.
//   tokens    = ids of run starts that are not blank          (predictions.py:198-199,206)
//   timesteps = 1-based frame index of each kept run's start  (predictions.py:201)
//   score     = sum_{t < len} max log-prob                    (predictions.py:206)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ctc_greedy_collapse_kernel(const int* __restrict__ argmax_in, const float* __restrict__ maxlp_in,
                                                                  const int* __restrict__ lengths, int n_utt, int T,
                                                                  int n_seq, int blank, int* __restrict__ tokens,
                                                                  int* __restrict__ timesteps, int* __restrict__ counts,
                                                                  float* __restrict__ scores) {
  const int seq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seq >= n_seq) return;
  const int lane = threadIdx.x & 31;
  const int utt = seq % n_utt;
  int len = lengths[utt];
  len = len < T ? (len < 0 ? 0 : len) : T;
  const int* ids = argmax_in + static_cast<long long>(seq) * T;
  const float* lp = maxlp_in + static_cast<long long>(seq) * T;
  int* tok = tokens + static_cast<long long>(seq) * T;
  int* ts = timesteps + static_cast<long long>(seq) * T;
  int count = 0;
  int carry = -1;  // id of frame t0 - 1
  float score = 0.f;
  for (int t0 = 0; t0 < len; t0 += 32) {
    const int t = t0 + lane;
    const bool in = t < len;
    const int id = in ? ids[t] : -1;
    if (in) score += lp[t];
    int prev = __shfl_up_sync(0xffffffffu, id, 1);
    if (lane == 0) prev = carry;
    const bool keep = in && (t == 0 || id != prev) && id != blank;
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = count + __popc(ballot & ((1u << lane) - 1u));
      tok[pos] = id;
      ts[pos] = t + 1;
    }
    count += __popc(ballot);
    carry = __shfl_sync(0xffffffffu, id, 31);
  }
  score = warp_sum(score);
  if (lane == 0) {
    counts[seq] = count;
    scores[seq] = score;
  }
}

// Packs the padded [n_seq][T] results into two dense arrays (hypothesis s occupies [offsets[s], offsets[s+1])):
// the host then splits one small tensor instead of mask-indexing n_seq * T elements.
// One block scans the counts (n_seq is heads * utterances, a few thousand); the copy then runs machine-wide, one warp per
// sequence (as a single block it took 97 us of a 17 ms end-to-end step).
__global__ void __launch_bounds__(256) ctc_pack_copy_kernel(const int* __restrict__ tokens, const int* __restrict__ timesteps,
                                                            const int* __restrict__ counts, int n_seq, int T,
                                                            const int* __restrict__ offsets, int* __restrict__ packed_tokens,
                                                            int* __restrict__ packed_timesteps) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= n_seq) return;
  const int off = offsets[s], c = counts[s];
  for (int i = lane; i < c; i += 32) {
    packed_tokens[off + i] = tokens[static_cast<long long>(s) * T + i];
    packed_timesteps[off + i] = timesteps[static_cast<long long>(s) * T + i];
  }
}

__global__ void __launch_bounds__(1024) ctc_pack_kernel(const int* __restrict__ tokens, const int* __restrict__ timesteps,
                                                        const int* __restrict__ counts, int n_seq, int T,
                                                        int* __restrict__ offsets /*[n_seq + 1]*/, int* __restrict__ packed_tokens,
                                                        int* __restrict__ packed_timesteps) {
  __shared__ int warp_totals[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < n_seq; s0 += 1024) {
    const int s = s0 + threadIdx.x;
    const int c = s < n_seq ? counts[s] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) warp_totals[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_totals[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += up;
      }
      warp_totals[lane] = w;  // inclusive scan of the warp totals
    }
    __syncthreads();
    const int before = base + (warp > 0 ? warp_totals[warp - 1] : 0) + inc - c;
    if (s < n_seq) offsets[s] = before;
    __syncthreads();
    if (threadIdx.x == 0) base += warp_totals[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n_seq] = base;
}

}  // namespace aph

using namespace aph;

extern "C" int aph_ctc_pack_hypotheses(const int32_t* tokens, const int32_t* timesteps, const int32_t* counts, int32_t n_seq,
                                       int32_t T, int32_t* offsets, int32_t* packed_tokens, int32_t* packed_timesteps,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(tokens && timesteps && counts && offsets && packed_tokens && packed_timesteps, "null pointer");
  APH_REQUIRE(n_seq > 0 && T > 0, "bad shape");
  ctc_pack_kernel<<<1, 1024, 0, stream>>>(tokens, timesteps, counts, n_seq, T, offsets, packed_tokens, packed_timesteps);
  ctc_pack_copy_kernel<<<static_cast<unsigned>((n_seq + 7) / 8), 256, 0, stream>>>(tokens, timesteps, counts, n_seq, T, offsets, packed_tokens,
                                                                                     packed_timesteps);
  APH_POST_LAUNCH(2);
  return APH_OK;
}

extern "C" int aph_compose_embeddings(const float* weight, int32_t n_categories, int32_t embedding_size, const int64_t* tfi,
                                      const int64_t* category_offsets, int32_t n_phonemes, int32_t n_features,
                                      int32_t rows_out, void* out_bf16, float* out_f32, int32_t* err_flag, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(weight && tfi && err_flag && (out_bf16 || out_f32), "null pointer");
  APH_REQUIRE(n_phonemes >= 0 && rows_out >= n_phonemes + 1 && embedding_size > 0 && n_features > 0, "bad shape");
  compose_embeddings_kernel<<<rows_out, 128, 0, stream>>>(weight, n_categories, embedding_size,
                                                          reinterpret_cast<const long long*>(tfi),
                                                          reinterpret_cast<const long long*>(category_offsets), n_phonemes,
                                                          n_features, rows_out, static_cast<__nv_bfloat16*>(out_bf16), out_f32,
                                                          err_flag);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_log_softmax_heads(const float* logits, int64_t ld, int64_t rows, int32_t col_lo, int32_t col_span,
                                     const int32_t* col_off, const int32_t* width, const int64_t* out_off, int32_t n_heads,
                                     float* out, int32_t* argmax_out, float* maxlp_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(logits && col_off && width && out_off && out, "null pointer");
  APH_REQUIRE(ld % 4 == 0 && col_lo % 4 == 0 && col_span % 4 == 0 && col_span > 0, "columns must be 16-byte aligned");
  APH_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "logits must be 16-byte aligned");
  if (rows <= 0 || n_heads <= 0) return APH_OK;
  const size_t smem = sizeof(float) * kLsmRows * static_cast<size_t>(col_span | 1);
  APH_REQUIRE(smem <= 200 * 1024, "packed narrow heads too wide for one tile; use aph_log_softmax_wide");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(log_softmax_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  const unsigned grid = static_cast<unsigned>((rows + kLsmRows - 1) / kLsmRows);
  log_softmax_heads_kernel<<<grid, 256, smem, stream>>>(logits, ld, rows, col_lo, col_span, col_off, width,
                                                       reinterpret_cast<const long long*>(out_off), n_heads, out, argmax_out,
                                                       maxlp_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_log_softmax_wide(const float* logits, int64_t ld, int64_t rows, int32_t width, float* out, int64_t ld_out,
                                    int32_t* argmax_out, float* maxlp_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(logits && out, "null pointer");
  APH_REQUIRE(width > 0, "bad width");
  if (rows <= 0) return APH_OK;
  const bool aligned = width % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const size_t bulk_smem = 128 + static_cast<size_t>(kLsmBulkWarps) * 2 * ((static_cast<size_t>(width) * 4 + 127) & ~static_cast<size_t>(127));
  if (aligned && width >= 256 && bulk_smem <= 110 * 1024) {
    static bool bulk_attr_set = false;
    if (!bulk_attr_set) {
      APH_CUDA_CHECK(cudaFuncSetAttribute(log_softmax_wide_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      bulk_attr_set = true;
    }
    long long blocks = (rows + kLsmBulkWarps - 1) / kLsmBulkWarps;
    const long long cap = 2LL * sm_count();
    if (blocks > cap) blocks = cap;
    log_softmax_wide_bulk_kernel<<<static_cast<unsigned>(blocks), kLsmBulkWarps * 32, bulk_smem, stream>>>(
        logits, ld, rows, width, out, ld_out, argmax_out, maxlp_out);
    APH_POST_LAUNCH(1);
    return APH_OK;
  }
  int warps = static_cast<int>((100 * 1024) / (sizeof(float) * width));
  APH_REQUIRE(warps >= 1, "row too wide for the shared-memory staged log_softmax");
  if (warps > 8) warps = 8;
  const size_t smem = sizeof(float) * static_cast<size_t>(warps) * width;
  static bool attr_set = false;
  if (!attr_set) {
    APH_CUDA_CHECK(cudaFuncSetAttribute(log_softmax_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  long long blocks = (rows + warps - 1) / warps;
  const long long cap = 2LL * sm_count() * 4;
  if (blocks > cap) blocks = cap;
  log_softmax_wide_kernel<<<static_cast<unsigned>(blocks), 256, smem, stream>>>(logits, ld, rows, width, out, ld_out, argmax_out,
                                                                              maxlp_out, warps);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

// ---------------------------------------------------------------------------
// Multi-head block kernels of the training step: every classifier head is a [rows][width] fp32 block inside some row-major
// matrix (its logits inside the level's output matrix, its gradient inside the level's gradient matrix, or a matrix of its
// own).  One launch walks all heads; the block descriptors travel by value in the kernel parameters.  They replace one torch
// launch PER HEAD each: the copies that hand the logits to autograd, the per-head log_softmax in front of the CTC loss and the
// strided adds that collect the logits gradients (3 x 37 launches of a few microseconds on the step's critical path).
// ---------------------------------------------------------------------------
constexpr int kMaxHeadBlocks = 48;
struct HeadBlockPack {
  aph_head_block src[kMaxHeadBlocks];
  aph_head_block dst[kMaxHeadBlocks];
};

__global__ void __launch_bounds__(256) copy_head_blocks_kernel(const __grid_constant__ HeadBlockPack pack, long long rows, int accumulate) {
  const aph_head_block src = pack.src[blockIdx.y];
  const aph_head_block dst = pack.dst[blockIdx.y];
  const float* in = static_cast<const float*>(src.ptr);
  float* out = static_cast<float*>(dst.ptr);
  const int width = src.width;
  const long long total = (src.rows > 0 ? src.rows : rows) * width;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / width;
    const int c = static_cast<int>(i - r * width);
    const float v = in[r * src.ld + c];
    float* o = out + r * dst.ld + c;
    *o = accumulate ? *o + v : v;
  }
}

// one warp per row: functional.log_softmax(x, -1) with the arithmetic of log_softmax_wide_kernel (expf / logf, fp32)
__global__ void __launch_bounds__(256) log_softmax_head_blocks_kernel(const __grid_constant__ HeadBlockPack pack, long long rows) {
  const aph_head_block src = pack.src[blockIdx.y];
  const aph_head_block dst = pack.dst[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int width = src.width;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  if (src.rows > 0) rows = src.rows;
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* in = static_cast<const float*>(src.ptr) + r * src.ld;
    float* out = static_cast<float*>(dst.ptr) + r * dst.ld;
    float m = -INFINITY;
    for (int c = lane; c < width; c += 32) m = fmaxf(m, in[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < width; c += 32) sum += expf(in[c] - m);
    sum = warp_sum(sum);
    const float lse = m + logf(sum);
    for (int c = lane; c < width; c += 32) out[c] = in[c] - lse;
  }
}

static int pack_head_blocks(const aph_head_block* src, const aph_head_block* dst, int first, int count, HeadBlockPack& pack) {
  for (int i = 0; i < count; ++i) {
    pack.src[i] = src[first + i];
    pack.dst[i] = dst[first + i];
    APH_REQUIRE(pack.src[i].ptr && pack.dst[i].ptr, "null block pointer");
    APH_REQUIRE(pack.src[i].width > 0 && pack.src[i].width == pack.dst[i].width, "source and destination blocks must have the same positive width");
    APH_REQUIRE(pack.src[i].ld >= pack.src[i].width && pack.dst[i].ld >= pack.dst[i].width, "row stride smaller than the block width");
    APH_REQUIRE(pack.src[i].rows >= 0 && pack.src[i].rows == pack.dst[i].rows, "source and destination blocks must have the same row count");
  }
  return APH_OK;
}

extern "C" int aph_copy_head_blocks(const aph_head_block* src_host, const aph_head_block* dst_host, int32_t n_blocks, int64_t rows,
                                    int32_t accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(src_host && dst_host, "null pointer");
  if (rows < 0 || n_blocks <= 0) return APH_OK;
  int launched = 0;
  for (int first = 0; first < n_blocks; first += kMaxHeadBlocks) {
    const int count = n_blocks - first < kMaxHeadBlocks ? n_blocks - first : kMaxHeadBlocks;
    HeadBlockPack pack;
    const int rc = pack_head_blocks(src_host, dst_host, first, count, pack);
    if (rc != APH_OK) return rc;
    long long largest = 0;
    for (int i = 0; i < count; ++i) {
      const long long elements = (pack.src[i].rows > 0 ? pack.src[i].rows : rows) * pack.src[i].width;
      largest = elements > largest ? elements : largest;
    }
    long long blocks = (largest + 1023) / 1024;
    if (blocks < 1) blocks = 1;
    if (blocks > 64) blocks = 64;
    copy_head_blocks_kernel<<<dim3(static_cast<unsigned>(blocks), count), 256, 0, stream>>>(pack, rows, accumulate);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_log_softmax_head_blocks(const aph_head_block* src_host, const aph_head_block* dst_host, int32_t n_blocks, int64_t rows,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(src_host && dst_host, "null pointer");
  if (rows <= 0 || n_blocks <= 0) return APH_OK;
  int launched = 0;
  for (int first = 0; first < n_blocks; first += kMaxHeadBlocks) {
    const int count = n_blocks - first < kMaxHeadBlocks ? n_blocks - first : kMaxHeadBlocks;
    HeadBlockPack pack;
    const int rc = pack_head_blocks(src_host, dst_host, first, count, pack);
    if (rc != APH_OK) return rc;
    long long blocks = (rows + 7) / 8;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    log_softmax_head_blocks_kernel<<<dim3(static_cast<unsigned>(blocks), count), 256, 0, stream>>>(pack, rows);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_dependency_softmax(const float* logits, int64_t ld, int64_t rows, const int32_t* col_off, const int32_t* width,
                                      const int32_t* dst_col, int32_t n_deps, int32_t skip, void* dst_bf16, int64_t ld_dst,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(logits && col_off && width && dst_col && dst_bf16, "null pointer");
  APH_REQUIRE(skip >= 0, "bad skip");
  if (rows <= 0 || n_deps <= 0) return APH_OK;
  long long blocks = (rows * n_deps + 255) / 256;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  dependency_softmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(logits, ld, rows, col_off, width, dst_col, n_deps,
                                                                              skip, static_cast<__nv_bfloat16*>(dst_bf16), ld_dst);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_argmax_rows(const float* x, int64_t ld, int64_t rows, int32_t width, int32_t* argmax_out, float* max_out,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(x && argmax_out && max_out, "null pointer");
  APH_REQUIRE(width > 0, "bad width");
  if (rows <= 0) return APH_OK;
  const long long per_block = width <= 32 ? 256 : 8;
  argmax_rows_kernel<<<static_cast<unsigned>((rows + per_block - 1) / per_block), 256, 0, stream>>>(x, ld, rows, width,
                                                                                                   argmax_out, max_out);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_ctc_greedy_collapse(const int32_t* argmax_in, const float* maxlp_in, const int32_t* lengths, int32_t n_utt,
                                       int32_t T, int32_t n_seq, int32_t blank, int32_t* tokens, int32_t* timesteps,
                                       int32_t* counts, float* scores, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(argmax_in && maxlp_in && lengths && tokens && timesteps && counts && scores, "null pointer");
  APH_REQUIRE(n_utt > 0 && n_seq % n_utt == 0 && T > 0, "n_seq must be heads * n_utt");
  ctc_greedy_collapse_kernel<<<ceil_div(n_seq, 4), 128, 0, stream>>>(argmax_in, maxlp_in, lengths, n_utt, T, n_seq, blank, tokens,
                                                                     timesteps, counts, scores);
  APH_POST_LAUNCH(1);
  return APH_OK;
}
