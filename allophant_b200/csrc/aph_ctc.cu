// allophant_b200 — multi-head CTC loss (forward alpha recursion, backward beta
// recursion fused with the gradient w.r.t. the LOGITS).
//
// Replaces, for ALL classifier heads of a step in one launch each way,
//   CTCWrapper.forward = nn.CTCLoss(reduction="sum", zero_infinity=True)(log_softmax(logits), ...)
//   (loss_functions.py:19-27, called per head at estimator.py:721-734 / 645-650)
// and its autograd backward (estimator.py:738).
//
// Two sets of recursion kernels.  While all (utterance, head) pairs fit the GPU at two blocks per SM (a training step:
// 37 heads x 8 utterances) a BLOCK owns a pair, one state per thread (ctc_pair_* further down); beyond that — and with
// APH_CTC_V1=1 — the warp kernels below:
// One warp owns one (utterance, head) pair and walks time sequentially in log
// space.  The 2S+1 CTC states are distributed over the lanes in contiguous
// chunks of K states (K = 2..32 by template), so only two values cross lanes
// per frame (warp shuffles).  Log-probs of narrow heads (<= 32 classes) are
// staged through shared memory 32 frames at a time with coalesced loads; wide
// heads gather only the label columns they need.  The blank index is 0
// (config.py:555 BLANK_OFFSET) and labels are the padded int64 [N, S_max]
// matrices of LabeledBatch (dataset_processing.py:132-162).
#include <string.h>

#include "aph_common.cuh"

namespace aph {

constexpr int kCtcWarps = 4;
constexpr int kCtcChunk = 32;    // frames staged per shared-memory refill
constexpr int kCtcSmallC = 32;   // heads up to this many classes use the staged path
constexpr int kCtcTinyC = 8;     // heads up to this many classes sum their occupancies per class in registers

// The recursions run in the LOG2 domain with the hardware approximations ex2.approx / lg2.approx (relative error
// 2^-22; the error of a 700-frame loss stays ~1e-7 relative, see test_ctc_long_labels_and_empty_targets) and without
// branches: the maximum is floored at a large finite value, so all-(-inf) inputs give lg2(0) = -inf by themselves.
// expf()/logf() with their range reduction and the `m == -inf` branches made the dependent chain 3x longer.
constexpr float kLog2E = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2a(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// The term of the maximum is exp2(0) = 1 and is not computed: lse2 costs 2 MUFU ops, lse3 3 (the recursions are bound by
// the MUFU issue rate of the ONE scheduler a warp lives on).  All-(-inf) inputs: the differences are taken against the
// floored maximum (finite), the result is built on the true one: -inf + lg2(1) = -inf.
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b), lo = fminf(a, b);
  const float mf = fmaxf(m, -1e30f);
  return m + lg2a(1.0f + ex2a(lo - mf));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  const float mf = fmaxf(m, -1e30f);
  return m + lg2a(1.0f + ex2a(mid - mf) + ex2a(lo - mf));
}

// One frame of the recursion for the K states of a lane:  out[i] = lse(cur[i], n1[i], n2[i] for odd i) + e[i]  (blank states,
// even i, have no skip transition; K is even, so the parity of i is the parity of the state).  The K evaluations are
// independent, but written one after the other the compiler also SCHEDULES them one after the other — max, subtract,
// ex2, add, lg2, add form a ~90-cycle dependent chain, K of them in series were ~900 cycles per frame on a warp that is
// alone on its scheduler (profiles/r02_ctc_recursion.md).  Here the work is laid out stage by stage over groups of up to 8
// states and the MUFU instructions are `asm volatile`, which keeps them in this order: a group's exponentials issue back to
// back, then its logarithms.
__device__ __forceinline__ float ex2v(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2v(float x) {
  float y;
  asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int K>
__device__ __forceinline__ void lse_frame(const float (&cur)[K], const float (&n1)[K], const float (&n2)[K], const float (&e)[K],
                                          float (&out)[K]) {
  constexpr int G = 8;
#pragma unroll
  for (int g0 = 0; g0 < K; g0 += G) {
    float m[G], d1[G], d2[G], x1[G], x2[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int i = g0 + j;
      if (i < K) {
        const float hi = fmaxf(cur[i], n1[i]), lo = fminf(cur[i], n1[i]);
        if (i & 1) {
          m[j] = fmaxf(hi, n2[i]);
          const float mid = fminf(hi, n2[i]);
          const float mf = fmaxf(m[j], -1e30f);
          d1[j] = mid - mf;
          d2[j] = lo - mf;
        } else {
          m[j] = hi;
          d1[j] = lo - fmaxf(hi, -1e30f);
          d2[j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < G; ++j)
      if (g0 + j < K) x1[j] = ex2v(d1[j]);
#pragma unroll
    for (int j = 1; j < G; j += 2)
      if (g0 + j < K) x2[j] = ex2v(d2[j]);
#pragma unroll
    for (int j = 0; j < G; ++j)
      if (g0 + j < K) x1[j] = (j & 1) ? (1.0f + x1[j]) + x2[j] : 1.0f + x1[j];
#pragma unroll
    for (int j = 0; j < G; ++j)
      if (g0 + j < K) x1[j] = lg2v(x1[j]);
#pragma unroll
    for (int j = 0; j < G; ++j)
      if (g0 + j < K) out[g0 + j] = (m[j] + x1[j]) + e[g0 + j];
  }
}

// The per-head descriptors travel BY VALUE in the kernel parameters (<= 4 KB): no device copy of the array, hence
// no synchronous pageable-memory transfer between the forward pass and the loss.
constexpr int kCtcMaxHeads = 48;
struct CtcHeadPack {
  aph_ctc_head h[kCtcMaxHeads];
};

struct PairInfo {
  const float* lp;      // log-probs of this utterance: element (t, k) at lp[t*stride_t + k]
  float* grad;          // same addressing, or NULL
  long long stride_t;
  int c;
  int T_in;             // valid frames
  int S;                // label length
  const int64_t* labels;
  float* alpha;         // [T][s_pad] workspace for this pair or NULL
  int s_pad;
};

__device__ __forceinline__ bool load_pair(const CtcHeadPack& heads, int h, int n, int n_utt, int T,
                                          const long long* input_lengths, float* alpha_ws, PairInfo& p) {
  const aph_ctc_head& hd = heads.h[h];
  p.lp = hd.log_probs + static_cast<long long>(n) * hd.stride_n;
  p.grad = hd.grad ? hd.grad + static_cast<long long>(n) * hd.stride_n : nullptr;
  p.stride_t = hd.stride_t;
  p.c = hd.n_classes;
  long long tl = input_lengths[n];
  p.T_in = static_cast<int>(tl < 0 ? 0 : (tl > T ? T : tl));
  long long sl = hd.label_lengths[n];
  p.S = static_cast<int>(sl < 0 ? 0 : sl);
  p.labels = hd.labels + static_cast<long long>(n) * hd.label_stride;
  p.s_pad = hd.s_pad;
  p.alpha = alpha_ws ? alpha_ws + hd.alpha_offset + static_cast<long long>(n) * T * hd.s_pad : nullptr;
  return true;
}

// ---------------------------------------------------------------------------
// forward: alpha recursion, per-pair negative log-likelihood
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kCtcWarps * 32) ctc_alpha_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads,
                                                                   int n_utt, int T,
                                                                   const long long* __restrict__ input_lengths,
                                                                   float* __restrict__ alpha_ws,
                                                                   float* __restrict__ nll_out /*[H][N]*/) {
  __shared__ float stage[kCtcWarps][kCtcChunk * kCtcSmallC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kCtcWarps + warp;
  if (pair >= n_heads * n_utt) return;
  const int h = pair / n_utt, n = pair - h * n_utt;
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, alpha_ws, p);
  const int S2 = 2 * p.S + 1;
  float* out = nll_out + static_cast<long long>(h) * n_utt + n;
  if (S2 > 32 * K || p.S > heads.h[h].label_stride || (p.alpha != nullptr && p.s_pad != 32 * K)) {  // host guarantees this never happens
    if (lane == 0) *out = NAN;
    return;
  }
  if (p.T_in == 0) {
    if (lane == 0) *out = p.S == 0 ? 0.f : INFINITY;
    return;
  }
  // per-lane state metadata
  int lab[K];
  bool skip[K];
  bool bad_label = false;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int s = lane * K + i;
    int l = 0;
    bool sk = false;
    if (s < S2 && (s & 1)) {
      l = static_cast<int>(p.labels[s >> 1]);
      if (s >= 3) sk = l != static_cast<int>(p.labels[(s >> 1) - 1]);
      if (l < 0 || l >= p.c) {  // a label outside the head's classes would index the staged log-probs out of bounds
        bad_label = true;
        l = 0;
      }
    }
    lab[i] = l;
    skip[i] = sk;
  }
  if (__any_sync(0xffffffffu, bad_label)) {  // nn.CTCLoss raises on such targets; here the loss (and its gradient) is NaN
    if (lane == 0) *out = NAN;
    return;
  }
  const bool small_c = p.c <= kCtcSmallC;
  float* st = stage[warp];
  float a[K];
  float nll = 0.f;
  bool poisoned = false;
  float e_next[K];
#pragma unroll
  for (int i = 0; i < K; ++i) e_next[i] = 0.f;
  // fmaxf / fminf treat a NaN as missing data, so a NaN frame (diverged logits: the whole log_softmax row is NaN) would be washed
  // out of the recursion from state 0 onwards: what is read from the log-probabilities is checked where it is loaded
  if (!small_c) {
    const float eb = __ldg(p.lp);
    poisoned = poisoned || (eb != eb);
#pragma unroll
    for (int i = 0; i < K; ++i) e_next[i] = (((lane * K + i) & 1) ? __ldg(p.lp + lab[i]) : eb) * kLog2E;
  }

  for (int t0 = 0; t0 < p.T_in; t0 += kCtcChunk) {
    const int nt = min(kCtcChunk, p.T_in - t0);
    if (small_c) {
      __syncwarp();
      const int total = nt * p.c;
      if (p.stride_t == p.c) {
        const float* src = p.lp + static_cast<long long>(t0) * p.stride_t;
        for (int i = lane; i < total; i += 32) {
          const float x = __ldg(src + i) * kLog2E;
          poisoned = poisoned || (x != x);
          st[i] = x;
        }
      } else {
        for (int i = lane; i < total; i += 32) {
          const int tt = i / p.c, k = i - tt * p.c;
          const float x = __ldg(p.lp + static_cast<long long>(t0 + tt) * p.stride_t + k) * kLog2E;
          poisoned = poisoned || (x != x);
          st[i] = x;
        }
      }
      __syncwarp();
    }
    for (int tt = 0; tt < nt; ++tt) {
      const int t = t0 + tt;
      float e[K];  // log2 emission probabilities
      if (small_c) {
#pragma unroll
        for (int i = 0; i < K; ++i) e[i] = st[tt * p.c + lab[i]];
      } else {
        // wide head: the gathers of frame t+1 are issued now and consumed one iteration later (a single warp per
        // scheduler cannot hide an L2 round trip behind anything else)
#pragma unroll
        for (int i = 0; i < K; ++i) e[i] = e_next[i];
        if (t + 1 < p.T_in) {
          const float* row = p.lp + static_cast<long long>(t + 1) * p.stride_t;
          const float eb = __ldg(row);
          poisoned = poisoned || (eb != eb);
#pragma unroll
          for (int i = 0; i < K; ++i) e_next[i] = (((lane * K + i) & 1) ? __ldg(row + lab[i]) : eb) * kLog2E;
        }
      }
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const int s = lane * K + i;
          a[i] = (s < 2 && s < S2) ? e[i] : -INFINITY;
        }
      } else {
        const float left1 = __shfl_up_sync(0xffffffffu, a[K - 1], 1);
        const float prev1 = lane == 0 ? -INFINITY : left1;  // state lane*K - 1: the only neighbour outside this lane's states
        float n1[K], n2[K], v[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          n1[i] = i == 0 ? prev1 : a[i - 1];
          // blank states (even s) have no skip transition
          n2[i] = ((i & 1) && skip[i]) ? (i == 1 ? prev1 : a[i - 2]) : -INFINITY;
        }
        lse_frame<K>(a, n1, n2, e, v);
#pragma unroll
        for (int i = 0; i < K; ++i) a[i] = (lane * K + i) < S2 ? v[i] : -INFINITY;
      }
      if (p.alpha) {
        // state lane*K + i is kept at column i*32 + lane of the row: each of the K stores covers 128 contiguous bytes
        float* dst = p.alpha + static_cast<long long>(t) * p.s_pad + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) dst[i * 32] = a[i];
      }
    }
  }
  // nll = -logsumexp(alpha_T-1(S2-1), alpha_T-1(S2-2))
  float last1 = -INFINITY, last2 = -INFINITY;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int s = lane * K + i;
    if (s == S2 - 1) last1 = a[i];
    if (s == S2 - 2) last2 = a[i];
  }
  // fetched from the lanes that own the two states (a maximum over the warp would drop a NaN: fmaxf(-inf, NaN) = -inf, and the
  // diverged loss would come out as +inf and be zeroed)
  last1 = __shfl_sync(0xffffffffu, last1, (S2 - 1) / K);
  last2 = S2 >= 2 ? __shfl_sync(0xffffffffu, last2, (S2 - 2) / K) : -INFINITY;
  nll = (__any_sync(0xffffffffu, poisoned) || last1 != last1 || last2 != last2) ? NAN : -lse2(last1, last2) * kLn2;  // natural units
  if (lane == 0) *out = nll;
}

// ---------------------------------------------------------------------------
// backward: beta recursion fused with d(sum of losses)/d(logits)
//   grad[t][k] = g * ( exp(lp[t][k]) - sum_{s: l'_s = k} exp(alpha_t(s) + beta_t(s) - lp[t][l'_s] + nll) )
// for t < T_in and finite nll; 0 otherwise (zero_infinity=True).
// Narrow heads: the class sums are accumulated in shared memory and the whole
// gradient row is written here.  Wide heads: the row exp(lp)*g was written by
// ctc_grad_init_kernel and the (few) label columns are corrected with atomics.
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kCtcWarps * 32) ctc_beta_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt,
                                                                  int T, const long long* __restrict__ input_lengths,
                                                                  const float* __restrict__ alpha_ws,
                                                                  const float* __restrict__ nll_in /*[H][N]*/,
                                                                  const float* __restrict__ grad_scale /*[H]*/) {
  __shared__ float stage[kCtcWarps][kCtcChunk * kCtcSmallC];
  __shared__ float csum[kCtcWarps][kCtcSmallC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kCtcWarps + warp;
  if (pair >= n_heads * n_utt) return;
  const int h = pair / n_utt, n = pair - h * n_utt;
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, const_cast<float*>(alpha_ws), p);
  if (p.grad == nullptr) return;
  const int S2 = 2 * p.S + 1;
  const float nll = nll_in[static_cast<long long>(h) * n_utt + n];
  const float nll2 = nll * kLog2E;  // the recursions (and the alpha workspace) are in the log2 domain
  const float g = grad_scale ? grad_scale[h] : 1.f;
  const bool small_c = p.c <= kCtcSmallC;
  // zero_infinity zeroes the gradient of an INFINITE loss only; a NaN loss (diverged logits, a label outside the classes)
  // poisons the gradient of its valid frames, as it does in torch
  const bool poisoned = nll != nll;
  const bool dead = nll == INFINITY || poisoned || S2 > 32 * K || p.s_pad != 32 * K;
  // frames past the utterance's length (and every frame of a zeroed loss) get zero gradient
  if (small_c) {
    const int t_zero_from = dead ? 0 : p.T_in;
    for (int t = t_zero_from; t < T; ++t)
      for (int k = lane; k < p.c; k += 32) p.grad[static_cast<long long>(t) * p.stride_t + k] = (poisoned && t < p.T_in) ? NAN : 0.f;
  }
  if (dead || p.T_in == 0) return;

  int lab[K];
  bool skip[K];  // transition s -> s+2 allowed
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int s = lane * K + i;
    int l = 0;
    bool sk = false;
    if (s < S2 && (s & 1)) {
      l = static_cast<int>(p.labels[s >> 1]);
      if (s + 2 < S2) sk = l != static_cast<int>(p.labels[(s >> 1) + 1]);
    }
    lab[i] = l;
    skip[i] = sk;
  }
  float* st = stage[warp];
  float* cs = csum[warp];
  float b[K];
  float al_next[K], e_next[K];
#pragma unroll
  for (int i = 0; i < K; ++i) al_next[i] = e_next[i] = 0.f;
  {  // column i*32 + lane of an alpha row holds state lane*K + i (see the alpha kernel)
    const float* nrow = p.alpha + static_cast<long long>(p.T_in - 1) * p.s_pad + lane;
#pragma unroll
    for (int i = 0; i < K; ++i) al_next[i] = nrow[i * 32];
  }
  if (!small_c) {
    const float* row = p.lp + static_cast<long long>(p.T_in - 1) * p.stride_t;
    const float eb = __ldg(row);
#pragma unroll
    for (int i = 0; i < K; ++i) e_next[i] = (((lane * K + i) & 1) ? __ldg(row + lab[i]) : eb) * kLog2E;
  }

  const int n_chunks = (p.T_in + kCtcChunk - 1) / kCtcChunk;
  for (int ch = n_chunks - 1; ch >= 0; --ch) {
    const int t0 = ch * kCtcChunk;
    const int nt = min(kCtcChunk, p.T_in - t0);
    if (small_c) {
      __syncwarp();
      const int total = nt * p.c;
      if (p.stride_t == p.c) {
        const float* src = p.lp + static_cast<long long>(t0) * p.stride_t;
        for (int i = lane; i < total; i += 32) st[i] = __ldg(src + i) * kLog2E;
      } else {
        for (int i = lane; i < total; i += 32) {
          const int tt = i / p.c, k = i - tt * p.c;
          st[i] = __ldg(p.lp + static_cast<long long>(t0 + tt) * p.stride_t + k) * kLog2E;
        }
      }
      __syncwarp();
    }
    for (int tt = nt - 1; tt >= 0; --tt) {
      const int t = t0 + tt;
      float e[K], al[K];
      // alpha row of this frame was requested one iteration ago; request the one of frame t-1 now
#pragma unroll
      for (int i = 0; i < K; ++i) al[i] = al_next[i];
      if (t > 0) {
        const float* nrow = p.alpha + static_cast<long long>(t - 1) * p.s_pad + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) al_next[i] = nrow[i * 32];
      }
      if (small_c) {
#pragma unroll
        for (int i = 0; i < K; ++i) e[i] = st[tt * p.c + lab[i]];
      } else {
#pragma unroll
        for (int i = 0; i < K; ++i) e[i] = e_next[i];
        if (t > 0) {
          const float* row = p.lp + static_cast<long long>(t - 1) * p.stride_t;
          const float eb = __ldg(row);
#pragma unroll
          for (int i = 0; i < K; ++i) e_next[i] = (((lane * K + i) & 1) ? __ldg(row + lab[i]) : eb) * kLog2E;
        }
      }
      if (t == p.T_in - 1) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const int s = lane * K + i;
          b[i] = (s < S2 && s >= S2 - 2) ? e[i] : -INFINITY;
        }
      } else {
        const float right1 = __shfl_down_sync(0xffffffffu, b[0], 1);
        const float right2 = __shfl_down_sync(0xffffffffu, b[1], 1);
        const float next1 = lane == 31 ? -INFINITY : right1;  // states (lane+1)*K and (lane+1)*K + 1
        const float next2 = lane == 31 ? -INFINITY : right2;
        float n1[K], n2[K], v[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          n1[i] = i == K - 1 ? next1 : b[i + 1];
          n2[i] = ((i & 1) && skip[i]) ? (i == K - 1 ? next2 : b[i + 2]) : -INFINITY;  // odd i <= K - 3, or the last state of the lane
        }
        lse_frame<K>(b, n1, n2, e, v);
        // states past 2S+1 stay at -inf by themselves: probability flows from s+1, s+2 to s, never out of the valid range
#pragma unroll
        for (int i = 0; i < K; ++i) b[i] = v[i];
      }
      // occupation probabilities gamma_t(s)
      float gam[K];
      float blank_sum = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const int s = lane * K + i;
        float gm = 0.f;
        if (s < S2) gm = ex2v(al[i] + b[i] - e[i] + nll2);  // ex2.approx.ftz: tiny occupancies flush to 0
        gam[i] = gm;
        if (!(s & 1)) blank_sum += gm;
      }
      blank_sum = warp_sum(blank_sum);
      if (small_c) {
        if (p.c <= kCtcTinyC) {
          // attribute heads (a handful of categories): per-class sums in registers + warp reductions; 32 lanes hammering 3-4
          // shared-memory addresses with atomics serialised the whole frame
          float acc[kCtcTinyC];
#pragma unroll
          for (int k = 0; k < kCtcTinyC; ++k) acc[k] = 0.f;
#pragma unroll
          for (int i = 1; i < K; i += 2) {  // label states (gam is 0 past S2)
#pragma unroll
            for (int k = 1; k < kCtcTinyC; ++k) acc[k] += lab[i] == k ? gam[i] : 0.f;
          }
#pragma unroll
          for (int k = 1; k < kCtcTinyC; ++k) acc[k] = warp_sum(acc[k]);
          if (lane < p.c) {
            float mine = 0.f;
#pragma unroll
            for (int k = 1; k < kCtcTinyC; ++k) mine = lane == k ? acc[k] : mine;
            cs[lane] = mine;
          }
        } else {
          if (lane < p.c) cs[lane] = 0.f;
          __syncwarp();
#pragma unroll
          for (int i = 0; i < K; ++i) {
            const int s = lane * K + i;
            if ((s & 1) && s < S2 && gam[i] != 0.f) atomicAdd(&cs[lab[i]], gam[i]);
          }
        }
        __syncwarp();
        if (lane < p.c) {
          const float occ = cs[lane] + (lane == 0 ? blank_sum : 0.f);
          p.grad[static_cast<long long>(t) * p.stride_t + lane] = g * (ex2a(st[tt * p.c + lane]) - occ);
        }
        __syncwarp();
      } else {
        float* grow = p.grad + static_cast<long long>(t) * p.stride_t;
        if (lane == 0) atomicAdd(grow, -g * blank_sum);
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const int s = lane * K + i;
          if ((s & 1) && s < S2 && gam[i] != 0.f) atomicAdd(grow + lab[i], -g * gam[i]);
        }
      }
    }
  }
}

// wide heads: grad[t][k] = g * exp(lp[t][k]) for valid frames of finite-loss pairs, else 0
__global__ void __launch_bounds__(256) ctc_grad_init_kernel(const aph_ctc_head hd, int h, int n_utt, int T,
                                                            const long long* __restrict__ input_lengths,
                                                            const float* __restrict__ nll_in,
                                                            const float* __restrict__ grad_scale) {
  const int n = blockIdx.y;
  const float nll = nll_in[static_cast<long long>(h) * n_utt + n];
  const float g = grad_scale ? grad_scale[h] : 1.f;
  long long tl = input_lengths[n];
  const bool poisoned = nll != nll;  // NaN loss: NaN gradient on the valid frames (the beta kernel leaves such pairs alone)
  const int t_valid = static_cast<int>(tl < 0 ? 0 : (tl > T ? T : tl));
  const int t_in = (nll < INFINITY) ? t_valid : 0;
  const long long total = static_cast<long long>(T) * hd.n_classes;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / hd.n_classes);
    const int k = static_cast<int>(i - static_cast<long long>(t) * hd.n_classes);
    const long long off = static_cast<long long>(n) * hd.stride_n + static_cast<long long>(t) * hd.stride_t + k;
    hd.grad[off] = t < t_in ? g * expf(hd.log_probs[off]) : ((poisoned && t < t_valid) ? NAN : 0.f);
  }
}

// per-head sum over the batch with zero_infinity (fixed order -> deterministic)
__global__ void ctc_reduce_kernel(const float* __restrict__ nll, int n_heads, int n_utt, float* __restrict__ loss_out) {
  const int h = blockIdx.x;
  const int lane = threadIdx.x;
  float s = 0.f;
  for (int n = lane; n < n_utt; n += 32) {
    const float v = nll[static_cast<long long>(h) * n_utt + n];
    if (v != INFINITY) s += v;  // zero_infinity drops +inf only: a NaN loss (diverged logits) stays visible, as in torch
  }
  s = warp_sum(s);
  if (lane == 0) loss_out[h] = s;
}

template <int K>
static int launch_alpha(const CtcHeadPack& heads, int n_heads, int n_utt, int T, const int64_t* input_lengths, float* alpha_ws,
                        float* nll, cudaStream_t stream) {
  const int pairs = n_heads * n_utt;
  ctc_alpha_kernel<K><<<ceil_div(pairs, kCtcWarps), kCtcWarps * 32, 0, stream>>>(
      heads, n_heads, n_utt, T, reinterpret_cast<const long long*>(input_lengths), alpha_ws, nll);
  return APH_OK;
}
template <int K>
static int launch_beta(const CtcHeadPack& heads, int n_heads, int n_utt, int T, const int64_t* input_lengths,
                       const float* alpha_ws, const float* nll, const float* grad_scale, cudaStream_t stream) {
  const int pairs = n_heads * n_utt;
  ctc_beta_kernel<K><<<ceil_div(pairs, kCtcWarps), kCtcWarps * 32, 0, stream>>>(
      heads, n_heads, n_utt, T, reinterpret_cast<const long long*>(input_lengths), alpha_ws, nll, grad_scale);
  return APH_OK;
}

// ---------------------------------------------------------------------------
// Long label sequences (more than 511 labels: 2S+1 states no longer fit 32 lanes x 32 registers).  nn.CTCLoss has no such
// limit (loss_functions.py:24); contour attributes emit several labels per phoneme and 30 s utterances can get there.  This is
// the plain formulation: one block per (utterance, head), the states strided over the threads, two alpha (beta) rows in shared
// memory, one block barrier per frame, natural-log arithmetic with expf / logf.  It is a correctness path — a batch whose longest
// label sequence stays within 511 never takes it.
// ---------------------------------------------------------------------------
constexpr int kCtcLongThreads = 256;

__device__ __forceinline__ float lse_nat(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m));
}

__global__ void __launch_bounds__(kCtcLongThreads) ctc_long_alpha_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt, int T,
                                                                         const long long* __restrict__ input_lengths,
                                                                         float* __restrict__ alpha_ws, float* __restrict__ nll_out) {
  extern __shared__ float long_smem[];  // [2][s_pad] rows, then int symbols[s_pad]
  const int pair = blockIdx.x;
  const int h = pair / n_utt, n = pair - h * n_utt;
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, alpha_ws, p);
  const int S2 = 2 * p.S + 1;
  float* out = nll_out + static_cast<long long>(h) * n_utt + n;
  if (S2 > p.s_pad || p.S > heads.h[h].label_stride) {
    if (threadIdx.x == 0) *out = NAN;
    return;
  }
  if (p.T_in == 0) {
    if (threadIdx.x == 0) *out = p.S == 0 ? 0.f : INFINITY;
    return;
  }
  float* row0 = long_smem;
  float* row1 = long_smem + p.s_pad;
  int* symbol = reinterpret_cast<int*>(long_smem + 2 * p.s_pad);
  for (int s = threadIdx.x; s < S2; s += blockDim.x) {
    symbol[s] = (s & 1) ? static_cast<int>(p.labels[s >> 1]) : 0;
    row0[s] = s < 2 ? p.lp[symbol[s]] : -INFINITY;
  }
  __syncthreads();
  if (p.alpha != nullptr)
    for (int s = threadIdx.x; s < S2; s += blockDim.x) p.alpha[s] = row0[s];
  float* prev = row0;
  float* cur = row1;
  for (int t = 1; t < p.T_in; ++t) {
    const float* lp_t = p.lp + static_cast<long long>(t) * p.stride_t;
    for (int s = threadIdx.x; s < S2; s += blockDim.x) {
      float v = prev[s];
      if (s > 0) v = lse_nat(v, prev[s - 1]);
      if ((s & 1) && s > 2 && symbol[s] != symbol[s - 2]) v = lse_nat(v, prev[s - 2]);
      v += lp_t[symbol[s]];
      cur[s] = v;
      if (p.alpha != nullptr) p.alpha[static_cast<long long>(t) * p.s_pad + s] = v;
    }
    __syncthreads();
    float* swap = prev;
    prev = cur;
    cur = swap;
  }
  if (threadIdx.x == 0) {
    float total = prev[S2 - 1];
    if (S2 > 1) total = lse_nat(total, prev[S2 - 2]);
    *out = -total;
  }
}

// beta recursion + gradient with respect to the logits behind the log_softmax: g * (exp(lp) - occupancy)
__global__ void __launch_bounds__(kCtcLongThreads) ctc_long_beta_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt, int T,
                                                                        const long long* __restrict__ input_lengths,
                                                                        const float* __restrict__ alpha_ws, const float* __restrict__ nll_in,
                                                                        const float* __restrict__ grad_scale) {
  extern __shared__ float long_smem[];
  const int pair = blockIdx.x;
  const int h = pair / n_utt, n = pair - h * n_utt;
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, const_cast<float*>(alpha_ws), p);
  if (p.grad == nullptr) return;
  const int S2 = 2 * p.S + 1;
  const float nll = nll_in[static_cast<long long>(h) * n_utt + n];
  const float g = grad_scale ? grad_scale[h] : 1.f;
  const bool dead = !(nll < INFINITY) || S2 > p.s_pad;
  const int t_valid = dead ? 0 : p.T_in;
  // frames past the utterance (all frames of a zeroed loss): zero gradient; valid frames start from g * softmax
  for (long long i = threadIdx.x; i < static_cast<long long>(T) * p.c; i += blockDim.x) {
    const int t = static_cast<int>(i / p.c);
    const int k = static_cast<int>(i - static_cast<long long>(t) * p.c);
    const long long off = static_cast<long long>(t) * p.stride_t + k;
    p.grad[off] = t < t_valid ? g * expf(p.lp[off]) : 0.f;
  }
  if (dead || p.T_in == 0) return;
  __syncthreads();
  float* row0 = long_smem;
  float* row1 = long_smem + p.s_pad;
  int* symbol = reinterpret_cast<int*>(long_smem + 2 * p.s_pad);
  const int t_last = p.T_in - 1;
  for (int s = threadIdx.x; s < S2; s += blockDim.x) {
    symbol[s] = (s & 1) ? static_cast<int>(p.labels[s >> 1]) : 0;
    row0[s] = s >= S2 - 2 ? p.lp[static_cast<long long>(t_last) * p.stride_t + symbol[s]] : -INFINITY;
  }
  __syncthreads();
  float* next = row0;
  float* cur = row1;
  for (int t = t_last; t >= 0; --t) {
    const float* lp_t = p.lp + static_cast<long long>(t) * p.stride_t;
    float* grad_t = p.grad + static_cast<long long>(t) * p.stride_t;
    const float* alpha_t = p.alpha + static_cast<long long>(t) * p.s_pad;
    if (t < t_last) {
      for (int s = threadIdx.x; s < S2; s += blockDim.x) {
        float v = next[s];
        if (s + 1 < S2) v = lse_nat(v, next[s + 1]);
        if ((s & 1) && s + 2 < S2 && symbol[s] != symbol[s + 2]) v = lse_nat(v, next[s + 2]);
        cur[s] = v + lp_t[symbol[s]];
      }
      __syncthreads();
      float* swap = next;
      next = cur;
      cur = swap;
    }
    // `next` now holds beta[t]
    for (int s = threadIdx.x; s < S2; s += blockDim.x) {
      const float lcab = alpha_t[s] + next[s];
      if (lcab > -INFINITY) atomicAdd(grad_t + symbol[s], -g * expf(lcab - lp_t[symbol[s]] + nll));
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Block-per-pair recursions (the default for label sequences up to 511): one block per (utterance, head), ONE STATE PER
// THREAD, two rows in shared memory, one block barrier per frame.  The warp-per-pair kernels above walk K = 10 states per
// lane in series on a warp that is alone on its scheduler (~1 000 cycles per frame, profiles/r02_ctc_recursion.md); here a
// frame is one log-sum-exp of three values per thread, ~10 warps per block and two to three blocks per SM.
//   alpha kernel: alpha rows [t][s] (log2 domain, natural state order) into the workspace, nll per pair;
//   beta kernel:  beta recursion; OVERWRITES alpha_t(s) with the state occupancy exp(alpha + beta - lp + nll);
//   grad kernel:  parallel over (pair, 48-frame chunk): the occupancies of a frame are summed per class (the states are
//                 visited in class order: a warp-segmented shuffle scan, the segment sums of a class added up by the thread at its first position)
//                 and the gradient row g * (softmax - occupancy sums) is written (narrow heads: the whole row; wide heads:
//                 the label columns of the row ctc_grad_init_kernel wrote).
// Nothing in the time loop of the two recursions depends on the number of classes.
// ---------------------------------------------------------------------------
constexpr int kCtcGradChunk = 48;    // frames per block of the gradient kernel (the class-order setup is O(states^2) per block)
constexpr int kCtcAhead = 8;  // frames per group of emission / alpha loads (even)

__device__ __forceinline__ int pair_symbol(const PairInfo& p, int s, int S2, bool& bad) {
  if (s >= S2 || !(s & 1)) return 0;
  int l = static_cast<int>(p.labels[s >> 1]);
  if (l < 0 || l >= p.c) {  // a label outside the head's classes: the loss is NaN (see the warp kernels)
    bad = true;
    l = 0;
  }
  return l;
}

// Emissions (and, backwards, alpha values) reach a thread kCtcAhead frames at a time: the loads of the NEXT group are issued
// at the top of a group and consumed a whole group (~2 800 cycles) later.  A rolling one-load-per-frame prefetch does not work:
// the wait on the oldest load is a wait on a scoreboard the newer loads share (ncu: a third of all warp samples on the first
// use of a value requested four frames earlier).
__global__ void __launch_bounds__(1024) ctc_pair_alpha_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt, int T,
                                                              const long long* __restrict__ input_lengths,
                                                              float* __restrict__ alpha_ws, float* __restrict__ nll_out) {
  extern __shared__ float pair_smem[];  // two rows of blockDim.x + 2 floats (two -inf guards in front of state 0)
  __shared__ int any_bad;
  const int pair = blockIdx.x;
  const int h = pair / n_utt, n = pair - h * n_utt;
  const int s = threadIdx.x;
  const int threads = static_cast<int>(blockDim.x);
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, alpha_ws, p);
  const int S2 = 2 * p.S + 1;
  float* out = nll_out + static_cast<long long>(h) * n_utt + n;
  if (S2 > threads || p.S > heads.h[h].label_stride || (p.alpha != nullptr && p.s_pad != threads)) {
    if (s == 0) *out = NAN;  // the host guarantees this never happens
    return;
  }
  if (p.T_in == 0) {
    if (s == 0) *out = p.S == 0 ? 0.f : INFINITY;
    return;
  }
  if (s == 0) any_bad = 0;
  __syncthreads();
  bool bad = false;
  const int sym = pair_symbol(p, s, S2, bad);
  const bool live = s < S2;
  const bool skip = (s & 1) && s >= 3 && live && sym != static_cast<int>(p.labels[(s >> 1) - 1]);  // transition s-2 -> s
  const int row_len = threads + 2;
  float* const row_even = pair_smem + 2;  // frame parity 0 / 1 (no pointer array: it would live in local memory)
  float* const row_odd = pair_smem + row_len + 2;
  if (s < 2) {
    pair_smem[s] = -INFINITY;
    pair_smem[row_len + s] = -INFINITY;
  }
  const long long stride_t = p.stride_t;
  const int T_in = p.T_in;
  const float* lp = p.lp + sym;
  float* alpha_out = (p.alpha != nullptr && live) ? p.alpha + s : nullptr;
  float a = (live && s < 2) ? __ldg(lp) * kLog2E : -INFINITY;
  bad = bad || a != a;
  row_even[s] = a;
  if (alpha_out != nullptr) *alpha_out = a;
  float e_next[kCtcAhead];
#pragma unroll
  for (int j = 0; j < kCtcAhead; ++j) e_next[j] = (live && 1 + j < T_in) ? __ldg(lp + static_cast<long long>(1 + j) * stride_t) : 0.f;
  __syncthreads();
  for (int tb = 1; tb < T_in; tb += kCtcAhead) {  // kCtcAhead is even: frame tb + j has the parity of 1 + j
    float e_cur[kCtcAhead];
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) e_cur[j] = e_next[j] * kLog2E;
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) {
      const int t = tb + kCtcAhead + j;
      e_next[j] = (live && t < T_in) ? __ldg(lp + static_cast<long long>(t) * stride_t) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) {
      const int t = tb + j;
      if (t < T_in) {  // block-uniform
        const float e = e_cur[j];
        bad = bad || (live && e != e);  // fmaxf drops NaNs: a diverged frame is remembered where its log-probabilities are used
        const float* prev = (j & 1) ? row_odd : row_even;  // frame t - 1
        float* cur = (j & 1) ? row_even : row_odd;
        const float a1 = prev[s - 1];
        const float a2 = skip ? prev[s - 2] : -INFINITY;
        a = live ? lse3(a, a1, a2) + e : -INFINITY;
        cur[s] = a;
        if (alpha_out != nullptr) alpha_out[static_cast<long long>(t) * threads] = a;
        __syncthreads();
      }
    }
  }
  if (bad) any_bad = 1;
  __syncthreads();
  if (s == 0) {
    const float* last = ((T_in - 1) & 1) ? row_odd : row_even;
    const float last1 = last[S2 - 1];
    const float last2 = S2 >= 2 ? last[S2 - 2] : -INFINITY;
    // (fmaxf would drop a NaN: the diverged loss must stay NaN, not become +inf and be zeroed)
    *out = (any_bad || last1 != last1 || last2 != last2) ? NAN : -lse2(last1, last2) * kLn2;  // natural units
  }
}

__global__ void __launch_bounds__(1024) ctc_pair_beta_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt, int T,
                                                             const long long* __restrict__ input_lengths, float* __restrict__ alpha_ws,
                                                             const float* __restrict__ nll_in) {
  extern __shared__ float pair_smem[];  // two rows of blockDim.x + 2 floats (two -inf guards behind the last state)
  const int pair = blockIdx.x;
  const int h = pair / n_utt, n = pair - h * n_utt;
  const int s = threadIdx.x;
  const int threads = static_cast<int>(blockDim.x);
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, alpha_ws, p);
  if (p.grad == nullptr) return;
  const int S2 = 2 * p.S + 1;
  const float nll = nll_in[static_cast<long long>(h) * n_utt + n];
  // an infinite loss has zero gradient (zero_infinity), a NaN loss a NaN gradient: the gradient kernel writes both
  if (!(nll < INFINITY) || S2 > threads || p.s_pad != threads || p.T_in == 0) return;
  const float nll2 = nll * kLog2E;
  bool bad = false;
  const int sym = pair_symbol(p, s, S2, bad);
  const bool live = s < S2;
  const bool skip = (s & 1) && s + 2 < S2 && sym != static_cast<int>(p.labels[(s >> 1) + 1]);  // transition s -> s+2
  const int row_len = threads + 2;
  if (s < 2) {
    pair_smem[threads + s] = -INFINITY;
    pair_smem[row_len + threads + s] = -INFINITY;
  }
  const long long stride_t = p.stride_t;
  const int t_last = p.T_in - 1;
  const float* lp = p.lp + sym;
  float* const alpha_io = p.alpha + s;
  float e_next[kCtcAhead], a_next[kCtcAhead];
#pragma unroll
  for (int j = 0; j < kCtcAhead; ++j) {
    const int t = t_last - j;
    e_next[j] = (live && t >= 0) ? __ldg(lp + static_cast<long long>(t) * stride_t) : 0.f;
    a_next[j] = (live && t >= 0) ? alpha_io[static_cast<long long>(t) * threads] : -INFINITY;
  }
  float b = -INFINITY;
  for (int tb = t_last; tb >= 0; tb -= kCtcAhead) {
    float e_cur[kCtcAhead], a_cur[kCtcAhead];
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) {
      e_cur[j] = e_next[j] * kLog2E;
      a_cur[j] = a_next[j];
    }
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) {
      const int t = tb - kCtcAhead - j;
      e_next[j] = (live && t >= 0) ? __ldg(lp + static_cast<long long>(t) * stride_t) : 0.f;
      a_next[j] = (live && t >= 0) ? alpha_io[static_cast<long long>(t) * threads] : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < kCtcAhead; ++j) {
      const int t = tb - j;
      if (t >= 0) {  // block-uniform
        const float e = e_cur[j];
        float* cur = pair_smem + (t & 1) * row_len;
        if (t == t_last) {
          b = (live && s >= S2 - 2) ? e : -INFINITY;
        } else {
          const float* next = pair_smem + ((t + 1) & 1) * row_len;
          const float b1 = next[s + 1];
          const float b2 = skip ? next[s + 2] : -INFINITY;
          b = live ? lse3(b, b1, b2) + e : -INFINITY;
        }
        cur[s] = b;
        // the occupancy of state s at frame t replaces alpha_t(s): exp(alpha + beta - lp + nll), 0 for unreachable states
        if (live) alpha_io[static_cast<long long>(t) * threads] = ex2a(a_cur[j] + b - e + nll2);
        __syncthreads();
      }
    }
  }
}

__global__ void __launch_bounds__(1024) ctc_pair_grad_kernel(const __grid_constant__ CtcHeadPack heads, int n_heads, int n_utt, int T,
                                                             const long long* __restrict__ input_lengths, const float* __restrict__ occ_ws,
                                                             const float* __restrict__ nll_in, const float* __restrict__ grad_scale) {
  extern __shared__ float pair_smem[];  // occupancies in class order [blockDim.x] | keys in class order [blockDim.x] | segment sums [32][32] | present [32]
  const int pair = blockIdx.y;
  const int h = pair / n_utt, n = pair - h * n_utt;
  const int tid = threadIdx.x, lane = tid & 31;
  const int threads = static_cast<int>(blockDim.x);
  PairInfo p;
  load_pair(heads, h, n, n_utt, T, input_lengths, const_cast<float*>(occ_ws), p);
  if (p.grad == nullptr) return;
  const int S2 = 2 * p.S + 1;
  const float nll = nll_in[static_cast<long long>(h) * n_utt + n];
  const float g = grad_scale ? grad_scale[h] : 1.f;
  const bool narrow = p.c <= kCtcSmallC;  // the whole gradient row is written here; wide rows start from ctc_grad_init_kernel
  const bool poisoned = nll != nll;
  const bool dead = !(nll < INFINITY) || S2 > threads || p.s_pad != threads;
  const int t0 = blockIdx.x * kCtcGradChunk;
  const int t1 = min(T, t0 + kCtcGradChunk);
  const int t_live = dead ? 0 : p.T_in;
  if (narrow) {  // frames past the utterance and every frame of a zeroed loss: zero gradient; NaN loss: NaN on its valid frames
    for (int t = max(t0, t_live); t < t1; ++t)
      if (tid < p.c) p.grad[static_cast<long long>(t) * p.stride_t + tid] = (poisoned && t < p.T_in) ? NAN : 0.f;
  }
  if (t0 >= t_live) return;
  float* occ_sm = pair_smem;
  int* key_sm = reinterpret_cast<int*>(pair_smem + threads);
  float* partial = pair_smem + 2 * threads;  // [warps][32]: the sum of every warp's i-th class segment
  int* present = reinterpret_cast<int*>(pair_smem + 2 * threads + 32 * 32);  // [32] narrow heads: class occurs among the states
  // states in class order: position = number of states with a smaller (class, state) key; padding threads sort last
  bool bad = false;
  const int key = tid < S2 ? pair_symbol(p, tid, S2, bad) : 0x7fffffff;
  key_sm[tid] = key;
  if (tid < 32) present[tid] = 0;
  __syncthreads();
  int pos = 0, same_key = 0;
  for (int j = 0; j < threads; j += 4) {  // threads is a multiple of 32
    const int4 other = *reinterpret_cast<const int4*>(key_sm + j);
    pos += (other.x < key || (other.x == key && j + 0 < tid)) ? 1 : 0;
    pos += (other.y < key || (other.y == key && j + 1 < tid)) ? 1 : 0;
    pos += (other.z < key || (other.z == key && j + 2 < tid)) ? 1 : 0;
    pos += (other.w < key || (other.w == key && j + 3 < tid)) ? 1 : 0;
    same_key += (other.x == key) + (other.y == key) + (other.z == key) + (other.w == key);
  }
  __syncthreads();
  key_sm[pos] = key;
  reinterpret_cast<int*>(occ_sm)[pos] = same_key;  // (scratch until the frame loop starts)
  __syncthreads();
  const int my_key = key_sm[tid];  // from here on the thread owns POSITION tid of the class order
  const int my_count = reinterpret_cast<int*>(occ_sm)[tid];
  const bool valid = my_key != 0x7fffffff;
  unsigned same = 0;  // bit d: the position 2^d to the left belongs to the same class and the same warp
#pragma unroll
  for (int d = 0; d < 5; ++d)
    if (lane >= (1 << d) && key_sm[tid - (1 << d)] == my_key) same |= 1u << d;
  const bool warp_tail = lane == 31 || tid + 1 >= threads || key_sm[tid + 1] != my_key;
  const bool class_head = valid && (tid == 0 || key_sm[tid - 1] != my_key);
  // No atomics: the tail of every class segment of a warp writes the segment's sum to partial[warp][segment index], and the
  // thread at a class's first position adds up the class's segments (its own warp's, then segment 0 of the warps the class
  // continues into).  All of this indexing is static over the frames.
  const unsigned tails = __ballot_sync(0xffffffffu, warp_tail);
  const int segment = __popc(tails & ((1u << lane) - 1u));
  const int first_warp = tid >> 5;
  const int last_warp = class_head ? (tid + my_count - 1) >> 5 : first_warp;
  if (narrow && class_head) present[my_key] = 1;
  __syncthreads();
  const bool absent_class = narrow && tid < p.c && present[tid] == 0;  // its gradient is g * softmax
  const int t_end = min(t1, t_live);
  const float* occ_row = p.alpha + static_cast<long long>(t0) * p.s_pad + tid;
  float occ_next = tid < S2 ? *occ_row : 0.f;
  for (int t = t0; t < t_end; ++t) {
    const float occ = occ_next;
    if (t + 1 < t_end && tid < S2) occ_next = occ_row[static_cast<long long>(t + 1 - t0) * p.s_pad];
    occ_sm[pos] = occ;
    __syncthreads();
    float v = occ_sm[tid];
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const float u = __shfl_up_sync(0xffffffffu, v, 1 << d);
      if (same & (1u << d)) v += u;
    }
    if (warp_tail) partial[first_warp * 32 + segment] = v;
    __syncthreads();
    float* grad_t = p.grad + static_cast<long long>(t) * p.stride_t;
    if (class_head) {
      float sum = partial[first_warp * 32 + segment];
      for (int w = first_warp + 1; w <= last_warp; ++w) sum += partial[w * 32];
      if (narrow)
        grad_t[my_key] = g * (ex2a(__ldg(p.lp + static_cast<long long>(t) * p.stride_t + my_key) * kLog2E) - sum);
      else
        grad_t[my_key] -= g * sum;
    }
    if (absent_class) grad_t[tid] = g * ex2a(__ldg(p.lp + static_cast<long long>(t) * p.stride_t + tid) * kLog2E);
    // (the next frame's first barrier orders these reads before the next writes of `partial`, and this frame's reads of occ_sm
    // before its writes)
  }
}

// Which recursion kernels a problem takes (the forward and the backward call must agree: the workspace layouts differ).  A
// block per pair cuts the latency of a frame in half, but spends ten times the instructions of the warp kernels on it: it is
// used while all pairs fit the GPU at two blocks per SM (the 37 heads x 8 utterances of a training step), the warp kernels
// beyond (64 utterances: 2 368 pairs are throughput-bound either way).  APH_CTC_V1=1 forces the warp kernels.
static int ctc_sm_count() {
  static int count = 0;
  if (count == 0) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess || cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) count = 1;
  }
  return count;
}
static bool ctc_use_warp_kernels(int pairs) {
  static int forced = -1;
  if (forced < 0) {
    const char* v = getenv("APH_CTC_V1");
    forced = (v != nullptr && v[0] == '1') ? 1 : 0;
  }
  return forced == 1 || pairs > 2 * ctc_sm_count();
}
// Dynamic shared memory of a recursion block: what it needs, padded so that no more than ceil(pairs / SMs) blocks fit one SM —
// otherwise the block scheduler stacks several blocks on some SMs while others stay empty, and a launch takes as long as
// its most crowded SM (measured: 296 pairs 880 cycles per frame unpadded against 350 for a single pair).
static size_t ctc_pair_smem(int pairs, size_t needed) {
  const int per_sm = (pairs + ctc_sm_count() - 1) / ctc_sm_count();
  const size_t share = static_cast<size_t>(220) * 1024 / static_cast<size_t>(per_sm < 1 ? 1 : per_sm);
  const size_t padded = share > 2048 ? share - 1024 : share;  // 1 KB per block is reserved by the driver
  return padded > needed ? padded : needed;
}

static size_t long_smem_bytes(int s_pad) { return static_cast<size_t>(s_pad) * 12; }

static int pick_k(int max_label_len) {
  const int states = 2 * max_label_len + 1;
  // even K only (the recursions specialise on the parity of the state); fine steps where the label lengths of speech live
  static const int kChoices[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32};
  for (int k : kChoices)
    if (states <= 32 * k) return k;
  return 0;
}

}  // namespace aph

using namespace aph;

// Longest label sequence the block-per-pair path takes: two rows of 2S+1 floats and the symbols in 200 KB of shared memory
constexpr int kCtcLongMaxLabels = 8000;

extern "C" int aph_ctc_states_pad(int32_t max_label_len) {
  const int k = pick_k(max_label_len);
  if (k != 0) return 32 * k;
  if (max_label_len > kCtcLongMaxLabels) return APH_ERR_UNSUPPORTED;
  return (2 * max_label_len + 1 + 31) / 32 * 32;  // block-per-pair path
}

extern "C" int aph_ctc_forward(const aph_ctc_head* heads_host, int32_t n_heads, int32_t n_utt, int32_t T,
                               int32_t max_label_len, const int64_t* input_lengths, float* alpha_ws, float* nll_out,
                               float* loss_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(heads_host && input_lengths && nll_out, "null pointer");
  APH_REQUIRE(n_heads > 0 && n_utt > 0 && T > 0, "empty problem");
  const int k = pick_k(max_label_len);
  if (k == 0 && max_label_len > kCtcLongMaxLabels) {
    set_last_error("aph_ctc_forward", "label sequences longer than 8000 are not supported", __FILE__, __LINE__);
    return APH_ERR_UNSUPPORTED;
  }
  int launched = 0;
  if (k == 0) {
    const size_t smem = long_smem_bytes(aph_ctc_states_pad(max_label_len));
    APH_CUDA_CHECK(cudaFuncSetAttribute(ctc_long_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  for (int h0 = 0; h0 < n_heads; h0 += kCtcMaxHeads) {
    const int nh = n_heads - h0 < kCtcMaxHeads ? n_heads - h0 : kCtcMaxHeads;
    CtcHeadPack pack;
    memcpy(pack.h, heads_host + h0, sizeof(aph_ctc_head) * nh);
    float* nll = nll_out + static_cast<long long>(h0) * n_utt;
    if (k == 0) {
      ctc_long_alpha_kernel<<<nh * n_utt, kCtcLongThreads, long_smem_bytes(aph_ctc_states_pad(max_label_len)), stream>>>(
          pack, nh, n_utt, T, reinterpret_cast<const long long*>(input_lengths), alpha_ws, nll);
      ++launched;
      continue;
    }
    if (!ctc_use_warp_kernels(n_heads * n_utt)) {
      const int s_pad = 32 * k;
      const size_t smem = ctc_pair_smem(nh * n_utt, 2 * (static_cast<size_t>(s_pad) + 2) * sizeof(float));
      APH_CUDA_CHECK(cudaFuncSetAttribute(ctc_pair_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      ctc_pair_alpha_kernel<<<nh * n_utt, s_pad, smem, stream>>>(pack, nh, n_utt, T, reinterpret_cast<const long long*>(input_lengths), alpha_ws, nll);
      ++launched;
      continue;
    }
    switch (k) {
      case 2: launch_alpha<2>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 4: launch_alpha<4>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 8: launch_alpha<8>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 6: launch_alpha<6>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 10: launch_alpha<10>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 12: launch_alpha<12>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 20: launch_alpha<20>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 24: launch_alpha<24>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      case 16: launch_alpha<16>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
      default: launch_alpha<32>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll, stream); break;
    }
    ++launched;
  }
  if (loss_out) {
    ctc_reduce_kernel<<<n_heads, 32, 0, stream>>>(nll_out, n_heads, n_utt, loss_out);
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}

extern "C" int aph_ctc_backward(const aph_ctc_head* heads_host, int32_t n_heads, int32_t n_utt, int32_t T,
                                int32_t max_label_len, const int64_t* input_lengths, const float* alpha_ws, const float* nll,
                                const float* grad_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_REQUIRE(heads_host && input_lengths && alpha_ws && nll, "null pointer");
  APH_REQUIRE(n_heads > 0 && n_utt > 0 && T > 0, "empty problem");
  const int k = pick_k(max_label_len);
  if (k == 0 && max_label_len > kCtcLongMaxLabels) {
    set_last_error("aph_ctc_backward", "label sequences longer than 8000 are not supported", __FILE__, __LINE__);
    return APH_ERR_UNSUPPORTED;
  }
  int launched = 0;
  if (k == 0) {  // block-per-pair path: initialises the gradient itself
    const size_t smem = long_smem_bytes(aph_ctc_states_pad(max_label_len));
    APH_CUDA_CHECK(cudaFuncSetAttribute(ctc_long_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    for (int h0 = 0; h0 < n_heads; h0 += kCtcMaxHeads) {
      const int nh = n_heads - h0 < kCtcMaxHeads ? n_heads - h0 : kCtcMaxHeads;
      CtcHeadPack pack;
      memcpy(pack.h, heads_host + h0, sizeof(aph_ctc_head) * nh);
      ctc_long_beta_kernel<<<nh * n_utt, kCtcLongThreads, smem, stream>>>(pack, nh, n_utt, T, reinterpret_cast<const long long*>(input_lengths),
                                                                         alpha_ws, nll + static_cast<long long>(h0) * n_utt,
                                                                         grad_scale ? grad_scale + h0 : nullptr);
      ++launched;
    }
    APH_POST_LAUNCH(launched);
    return APH_OK;
  }
  for (int h = 0; h < n_heads; ++h) {
    if (heads_host[h].n_classes > kCtcSmallC && heads_host[h].grad != nullptr) {
      long long blocks = (static_cast<long long>(T) * heads_host[h].n_classes + 255) / 256;
      if (blocks > 64) blocks = 64;
      ctc_grad_init_kernel<<<dim3(static_cast<unsigned>(blocks), n_utt), 256, 0, stream>>>(
          heads_host[h], h, n_utt, T, reinterpret_cast<const long long*>(input_lengths), nll, grad_scale);
      ++launched;
    }
  }
  for (int h0 = 0; h0 < n_heads; h0 += kCtcMaxHeads) {
    const int nh = n_heads - h0 < kCtcMaxHeads ? n_heads - h0 : kCtcMaxHeads;
    CtcHeadPack pack;
    memcpy(pack.h, heads_host + h0, sizeof(aph_ctc_head) * nh);
    const float* nll_h = nll + static_cast<long long>(h0) * n_utt;
    const float* scale_h = grad_scale ? grad_scale + h0 : nullptr;
    if (!ctc_use_warp_kernels(n_heads * n_utt)) {
      const int s_pad = 32 * k;
      const size_t smem = ctc_pair_smem(nh * n_utt, 2 * (static_cast<size_t>(s_pad) + 2) * sizeof(float));
      APH_CUDA_CHECK(cudaFuncSetAttribute(ctc_pair_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      const size_t grad_smem = (2 * static_cast<size_t>(s_pad) + 32 * 32 + 32) * sizeof(float);
      if (grad_smem > 48 * 1024)
        APH_CUDA_CHECK(cudaFuncSetAttribute(ctc_pair_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(grad_smem)));
      ctc_pair_beta_kernel<<<nh * n_utt, s_pad, smem, stream>>>(pack, nh, n_utt, T, reinterpret_cast<const long long*>(input_lengths),
                                                                const_cast<float*>(alpha_ws), nll_h);
      ctc_pair_grad_kernel<<<dim3(ceil_div(T, kCtcGradChunk), nh * n_utt), s_pad, grad_smem, stream>>>(
          pack, nh, n_utt, T, reinterpret_cast<const long long*>(input_lengths), alpha_ws, nll_h, scale_h);
      launched += 2;
      continue;
    }
    switch (k) {
      case 2: launch_beta<2>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 4: launch_beta<4>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 8: launch_beta<8>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 6: launch_beta<6>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 10: launch_beta<10>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 12: launch_beta<12>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 20: launch_beta<20>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 24: launch_beta<24>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      case 16: launch_beta<16>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
      default: launch_beta<32>(pack, nh, n_utt, T, input_lengths, alpha_ws, nll_h, scale_h, stream); break;
    }
    ++launched;
  }
  APH_POST_LAUNCH(launched);
  return APH_OK;
}
