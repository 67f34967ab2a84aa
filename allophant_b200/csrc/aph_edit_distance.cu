// allophant_b200 — host-side edit distance (PER / AER bookkeeping), C++ replacement of the reference's
// Rust extension `allophant.phonemes` (src/edit_distance.rs: levensthein 70-96,
// levensthein_statistics 601-608 -> _general 372-481 with uniform_costs 483-496).
//
// The reference evaluates one pair of Python lists per call.  Here a whole evaluation batch
// (all utterances x all classifiers) is one call over flat int64 symbol arrays, spread over host
// threads; each worker keeps one reusable cost matrix.  Results must be IDENTICAL to the Rust
// code, including its tie-breaking (deletion only if strictly cheaper than insertion; diagonal
// move when it is <= that; "correct" when the diagonal cost equals the current cost) and its f32
// cost arithmetic, because PER/AER are compared for equality.
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "aph_common.cuh"

namespace aph {

struct EditScratch {
  std::vector<float> cost;  // (m+1) x (n+1)
};

static void edit_statistics_one(const int64_t* a, int64_t m, const int64_t* b, int64_t n, EditScratch& scratch, uint64_t out[4]) {
  const int64_t w = n + 1;
  scratch.cost.resize(static_cast<size_t>((m + 1) * w));
  float* cost = scratch.cost.data();
  for (int64_t j = 0; j <= n; ++j) cost[j] = static_cast<float>(j);
  for (int64_t i = 1; i <= m; ++i) {
    const float* up = cost + (i - 1) * w;
    float* row = cost + i * w;
    row[0] = up[0] + 1.0f;
    const int64_t symbol = a[i - 1];
    for (int64_t j = 1; j <= n; ++j) {
      const float deletion = up[j] + 1.0f;
      const float insertion = row[j - 1] + 1.0f;
      const float substitution = up[j - 1] + (symbol != b[j - 1] ? 1.0f : 0.0f);
      row[j] = std::min(std::min(insertion, deletion), substitution);
    }
  }
  uint64_t insertions = 0, deletions = 0, substitutions = 0, correct = 0;
  int64_t i = m, j = n;
  float current = cost[m * w + n];
  while (current != 0.0f) {
    if (i == 0) {
      if (j == 0) break;
      current = cost[j - 1];
      --j;
      ++insertions;
      continue;
    }
    if (j == 0) {
      current = cost[(i - 1) * w];
      --i;
      ++deletions;
      continue;
    }
    const float deletion = cost[(i - 1) * w + j];
    const float insertion = cost[i * w + j - 1];
    const float diagonal = cost[(i - 1) * w + j - 1];
    const bool take_deletion = deletion < insertion;
    const float side = take_deletion ? deletion : insertion;
    if (diagonal <= side) {
      if (diagonal == current) {
        ++correct;
      } else {
        ++substitutions;
      }
      current = diagonal;
      --i;
      --j;
    } else if (take_deletion) {
      current = deletion;
      --i;
      ++deletions;
    } else {
      current = insertion;
      --j;
      ++insertions;
    }
  }
  out[0] = insertions;
  out[1] = deletions;
  out[2] = substitutions;
  out[3] = correct + static_cast<uint64_t>(i);  // remaining prefix of the expected sequence counts as correct
}

static uint64_t edit_distance_one(const int64_t* a, int64_t m, const int64_t* b, int64_t n, std::vector<uint64_t>& rows) {
  rows.resize(static_cast<size_t>(2 * (n + 1)));
  uint64_t* previous = rows.data();
  uint64_t* current = previous + (n + 1);
  for (int64_t j = 0; j <= n; ++j) previous[j] = static_cast<uint64_t>(j);
  for (int64_t i = 0; i < m; ++i) {
    current[0] = static_cast<uint64_t>(i) + 1;
    for (int64_t j = 0; j < n; ++j) {
      const uint64_t deletion = previous[j + 1] + 1;
      const uint64_t insertion = current[j] + 1;
      const uint64_t substitution = previous[j] + (a[i] != b[j] ? 1u : 0u);
      current[j + 1] = std::min(std::min(deletion, insertion), substitution);
    }
    std::swap(previous, current);
  }
  return previous[n];
}

}  // namespace aph

using namespace aph;

// Pair p compares expected[expected_offsets[p] : expected_offsets[p+1]] with
// actual[actual_offsets[p] : actual_offsets[p+1]].  All pointers are HOST pointers.
extern "C" int aph_edit_statistics_batch(const int64_t* expected_host, const int64_t* expected_offsets_host,
                                         const int64_t* actual_host, const int64_t* actual_offsets_host, int64_t n_pairs,
                                         uint64_t* statistics_host /*[n_pairs][4]: I, D, S, C*/,
                                         uint64_t* distances_host /*[n_pairs] or NULL*/, int32_t n_threads) {
  APH_REQUIRE(expected_offsets_host && actual_offsets_host && (statistics_host || distances_host), "null pointer");
  APH_REQUIRE(n_pairs >= 0, "negative pair count");
  if (n_pairs == 0) return APH_OK;
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  threads = std::max(1, std::min<int>(threads, static_cast<int>(std::min<int64_t>(n_pairs, 256))));
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    EditScratch scratch;
    std::vector<uint64_t> rows;
    for (;;) {
      const int64_t p = next.fetch_add(1);
      if (p >= n_pairs) break;
      const int64_t* a = expected_host + expected_offsets_host[p];
      const int64_t m = expected_offsets_host[p + 1] - expected_offsets_host[p];
      const int64_t* b = actual_host + actual_offsets_host[p];
      const int64_t n = actual_offsets_host[p + 1] - actual_offsets_host[p];
      if (statistics_host) edit_statistics_one(a, m, b, n, scratch, statistics_host + 4 * p);
      if (distances_host) distances_host[p] = edit_distance_one(a, m, b, n, rows);
    }
  };
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
  }
  return APH_OK;
}

// EditStatistics::word_error_rate (src/edit_distance.rs:311-317): (S + D + I) / (S + D + C) in f32.
extern "C" float aph_word_error_rate(uint64_t insertions, uint64_t deletions, uint64_t substitutions, uint64_t correct) {
  const float substituted_or_deleted = static_cast<float>(substitutions + deletions);
  return (substituted_or_deleted + static_cast<float>(insertions)) / (substituted_or_deleted + static_cast<float>(correct));
}

// ---- host-side collation (batching.py:171-215, rnn.pad_sequence of the audio) -------------------------------------
// Copies n variable-length fp32 utterances into one zero-padded [n][max_len] matrix (normally a pinned staging buffer
// that the host -> device copy then reads), spread over host threads.  ALL POINTERS ARE HOST POINTERS.
#include <string.h>

#include <thread>
#include <vector>

extern "C" int aph_collate_pad_f32(const float* const* utterances_host, const int64_t* lengths_host, int64_t n, int64_t max_len,
                                   float* dst_host, int32_t n_threads) {
  if (!utterances_host || !lengths_host || !dst_host || n < 0 || max_len < 0) return APH_ERR_INVALID;
  for (int64_t i = 0; i < n; ++i)
    if (lengths_host[i] < 0 || lengths_host[i] > max_len || (lengths_host[i] > 0 && utterances_host[i] == nullptr)) return APH_ERR_INVALID;
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (threads < 1) threads = 1;
  if (threads > n) threads = static_cast<int>(n > 0 ? n : 1);
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      float* row = dst_host + i * max_len;
      const int64_t len = lengths_host[i];
      if (len > 0) memcpy(row, utterances_host[i], sizeof(float) * static_cast<size_t>(len));
      if (len < max_len) memset(row + len, 0, sizeof(float) * static_cast<size_t>(max_len - len));
    }
  };
  if (threads == 1 || n * max_len < (1 << 18)) {
    work(0, n);
    return APH_OK;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return APH_OK;
}
