// Train-mode stochastic regularisation of the wav2vec2 encoder (Hugging Face modeling_wav2vec2.py: feature projection
// dropout 431-433, SpecAugment _compute_mask_indices / _mask_hidden_states, encoder dropout 766, LayerDrop 774-777,
// attention dropout, hidden dropout of the layers 742/752) and of the classifier inputs (acoustic_model.py:486-488).
//
// All masks are COUNTER-BASED (aph_common.cuh: drop_hash): the forward kernels (GEMM epilogue, attention, the kernels
// here) and the backward kernels regenerate the same keep decision from (seed, row, column); no mask is ever stored.
// The elementwise kernels here are HBM-bound: one read + one write of the activation (8 or 6 B per element).
#include "aph_common.cuh"

namespace aph {

// out = x * keep * scale; rows in row_mask (SpecAugment, optional) are replaced by row_fill.  In place when out_f32 == x.
__global__ void __launch_bounds__(256) dropout_2d_kernel(const float* __restrict__ x, long long ld_x, long long rows, int cols,
                                                         uint32_t threshold, uint32_t seed, float scale,
                                                         const uint8_t* __restrict__ row_mask, const float* __restrict__ row_fill,
                                                         float* out_f32, long long ld_f32, __nv_bfloat16* out_bf16, long long ld_bf16) {
  const int quads = cols >> 2;
  const long long total = rows * quads;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / quads;
    const int col = static_cast<int>(i - row * quads) << 2;
    float4 v = *reinterpret_cast<const float4*>(x + row * ld_x + col);
    if (row_mask != nullptr && row_mask[row]) {
      v = row_fill != nullptr ? *reinterpret_cast<const float4*>(row_fill + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (threshold != 0) {
      const uint32_t key = drop_row_key(seed, static_cast<uint32_t>(row));
      const uint32_t h0 = drop_hash(key, static_cast<uint32_t>(col >> 1));
      const uint32_t h1 = drop_hash(key, static_cast<uint32_t>(col >> 1) + 1u);
      v.x = drop_keep(h0, 0, threshold) ? v.x * scale : 0.f;
      v.y = drop_keep(h0, 1, threshold) ? v.y * scale : 0.f;
      v.z = drop_keep(h1, 0, threshold) ? v.z * scale : 0.f;
      v.w = drop_keep(h1, 1, threshold) ? v.w * scale : 0.f;
    }
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + row * ld_f32 + col) = v;
    if (out_bf16 != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out_bf16 + row * ld_bf16 + col) = o;
    }
  }
}

// bf16 in place (the classifier feature matrix X): x = x * keep * scale
__global__ void __launch_bounds__(256) dropout_bf16_2d_kernel(__nv_bfloat16* x, long long ld, long long rows, int cols, uint32_t threshold,
                                                              uint32_t seed, float scale) {
  const int quads = cols >> 2;
  const long long total = rows * quads;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / quads;
    const int col = static_cast<int>(i - row * quads) << 2;
    uint2* ptr = reinterpret_cast<uint2*>(x + row * ld + col);
    const uint2 raw = *ptr;
    float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
    const uint32_t key = drop_row_key(seed, static_cast<uint32_t>(row));
    const uint32_t h0 = drop_hash(key, static_cast<uint32_t>(col >> 1));
    const uint32_t h1 = drop_hash(key, static_cast<uint32_t>(col >> 1) + 1u);
    a.x = drop_keep(h0, 0, threshold) ? a.x * scale : 0.f;
    a.y = drop_keep(h0, 1, threshold) ? a.y * scale : 0.f;
    b.x = drop_keep(h1, 0, threshold) ? b.x * scale : 0.f;
    b.y = drop_keep(h1, 1, threshold) ? b.y * scale : 0.f;
    uint2 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(b.x, b.y);
    *ptr = o;
  }
}

// SpecAugment time mask (HF _compute_mask_indices): per utterance, spans = max(int(prob * len / span + eps), min_masks),
// clipped to what fits, span starts drawn WITHOUT replacement from [0, len - span + 1); spans may overlap.  eps is one
// uniform draw shared by the batch, as in HF.  One block per utterance.
__global__ void __launch_bounds__(128) spec_augment_mask_kernel(const int* __restrict__ frames, int seq, float prob, int span, int min_masks,
                                                                uint32_t seed, uint8_t* __restrict__ mask) {
  constexpr int kMaxSpans = 512;
  __shared__ int starts[kMaxSpans];
  __shared__ int n_spans;
  const int utt = blockIdx.x;
  uint8_t* row = mask + static_cast<long long>(utt) * seq;
  for (int t = threadIdx.x; t < seq; t += blockDim.x) row[t] = 0;
  if (threadIdx.x == 0) {
    const int len = min(frames[utt], seq);
    const float eps = static_cast<float>(drop_fmix(seed ^ 0xA5A5A5A5u) >> 8) * (1.0f / 16777216.0f);
    int count = static_cast<int>(prob * static_cast<float>(len) / static_cast<float>(span) + eps);
    count = max(count, min_masks);
    if (count * span > seq) count = seq / span;
    const int choices = len - (span - 1);
    if (choices < count) count = max(choices, 0);
    count = min(count, kMaxSpans);
    const uint32_t key = drop_row_key(seed, static_cast<uint32_t>(utt) + 1u);
    uint32_t draw = 0;
    for (int i = 0; i < count; ++i) {
      int start;
      bool fresh;
      do {  // rejection of repeated starts: count <= choices, so this terminates
        start = static_cast<int>(drop_hash(key, draw++) % static_cast<uint32_t>(choices));
        fresh = true;
        for (int j = 0; j < i; ++j) fresh = fresh && starts[j] != start;
      } while (!fresh);
      starts[i] = start;
    }
    n_spans = count;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_spans * span; i += blockDim.x) {
    const int t = min(starts[i / span] + i % span, seq - 1);
    row[t] = 1;
  }
}

// SpecAugment along the feature axis (HF _mask_hidden_states, mask_feature_prob): x[n][t][c] = 0 where col_mask[n][c];
// fp32 in place + the optional bf16 copy.  The same call on a gradient is the backward pass.
__global__ void __launch_bounds__(256) mask_columns_kernel(float* x, long long ld, long long rows, int seq, int cols,
                                                           const uint8_t* __restrict__ col_mask, __nv_bfloat16* x_bf16, long long ld_bf16) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long row = i / cols;
    const int c = static_cast<int>(i - row * cols);
    if (col_mask[(row / seq) * cols + c]) {
      x[row * ld + c] = 0.f;
      if (x_bf16 != nullptr) x_bf16[row * ld_bf16 + c] = __float2bfloat16(0.f);
    }
  }
}

// backward of the masked rows: d_fill[col] += sum over masked rows of d[row][col]; those rows of d become 0.
__global__ void __launch_bounds__(256) masked_rows_backward_kernel(float* d, long long ld, long long rows, int cols,
                                                                   const uint8_t* __restrict__ row_mask, float* __restrict__ d_fill) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= cols) return;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long lo = blockIdx.y * per, hi = min(rows, lo + per);
  float sum = 0.f;
  for (long long r = lo; r < hi; ++r) {
    if (row_mask[r]) {
      sum += d[r * ld + col];
      d[r * ld + col] = 0.f;
    }
  }
  if (sum != 0.f) atomicAdd(d_fill + col, sum);
}

}  // namespace aph

using namespace aph;

static unsigned grid_for(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = 148ll * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

extern "C" int aph_dropout_2d(const float* x, int64_t ld_x, int64_t rows, int32_t cols, uint32_t threshold, uint32_t seed, float scale,
                              const uint8_t* row_mask, const float* row_fill, float* out_f32, int64_t ld_f32, void* out_bf16,
                              int64_t ld_bf16, void* stream_) {
  APH_REQUIRE(x && rows >= 0 && cols > 0 && cols % 4 == 0 && ld_x % 4 == 0, "dropout_2d: columns and ld must be multiples of 4");
  APH_REQUIRE(out_f32 || out_bf16, "dropout_2d: no output");
  APH_REQUIRE(!out_f32 || ld_f32 % 4 == 0, "dropout_2d: ld_f32 % 4");
  APH_REQUIRE(!out_bf16 || ld_bf16 % 4 == 0, "dropout_2d: ld_bf16 % 4");
  APH_REQUIRE(threshold < 65536u, "dropout_2d: threshold is 16 bits (p < 1)");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dropout_2d_kernel<<<grid_for(rows * (cols / 4)), 256, 0, stream>>>(x, ld_x, rows, cols, threshold, seed, scale, row_mask, row_fill, out_f32,
                                                                      ld_f32, static_cast<__nv_bfloat16*>(out_bf16), ld_bf16);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_dropout_bf16_2d(void* x_bf16, int64_t ld, int64_t rows, int32_t cols, uint32_t threshold, uint32_t seed, float scale,
                                   void* stream_) {
  APH_REQUIRE(x_bf16 && rows >= 0 && cols > 0 && cols % 4 == 0 && ld % 4 == 0, "dropout_bf16_2d: columns and ld must be multiples of 4");
  APH_REQUIRE(threshold < 65536u, "dropout_bf16_2d: threshold is 16 bits (p < 1)");
  if (rows == 0 || threshold == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dropout_bf16_2d_kernel<<<grid_for(rows * (cols / 4)), 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x_bf16), ld, rows, cols, threshold, seed,
                                                                           scale);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_spec_augment_mask(const int32_t* frames, int32_t n_utt, int32_t seq, float mask_prob, int32_t mask_length,
                                     int32_t min_masks, uint32_t seed, uint8_t* mask, void* stream_) {
  APH_REQUIRE(frames && mask && n_utt >= 0 && seq > 0, "spec_augment_mask: bad arguments");
  APH_REQUIRE(mask_length >= 1, "spec_augment_mask: `mask_length` has to be bigger than 0.");
  APH_REQUIRE(mask_length <= seq, "spec_augment_mask: `mask_length` has to be smaller than `sequence_length`");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  spec_augment_mask_kernel<<<static_cast<unsigned>(n_utt), 128, 0, stream>>>(frames, seq, mask_prob, mask_length, min_masks, seed, mask);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_masked_rows_backward(float* d, int64_t ld, int64_t rows, int32_t cols, const uint8_t* row_mask, float* d_fill,
                                        void* stream_) {
  APH_REQUIRE(d && row_mask && d_fill && rows >= 0 && cols > 0, "masked_rows_backward: bad arguments");
  if (rows == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  APH_CUDA_CHECK(cudaMemsetAsync(d_fill, 0, sizeof(float) * static_cast<size_t>(cols), stream));
  dim3 grid(static_cast<unsigned>((cols + 255) / 256), static_cast<unsigned>(rows < 64 ? 1 : (rows / 64 > 296 ? 296 : rows / 64)));
  masked_rows_backward_kernel<<<grid, 256, 0, stream>>>(d, ld, rows, cols, row_mask, d_fill);
  APH_POST_LAUNCH(1);
  return APH_OK;
}

extern "C" int aph_mask_columns(float* x, int64_t ld, int32_t n_utt, int32_t seq, int32_t cols, const uint8_t* col_mask, void* x_bf16,
                                int64_t ld_bf16, void* stream_) {
  APH_REQUIRE(x && col_mask && n_utt >= 0 && seq > 0 && cols > 0, "mask_columns: bad arguments");
  if (n_utt == 0) return APH_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long rows = static_cast<long long>(n_utt) * seq;
  mask_columns_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(x, ld, rows, seq, cols, col_mask, static_cast<__nv_bfloat16*>(x_bf16), ld_bf16);
  APH_POST_LAUNCH(1);
  return APH_OK;
}
