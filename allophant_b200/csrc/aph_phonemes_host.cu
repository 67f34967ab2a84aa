// allophant_b200 — host-side remainder of the reference's Rust extension `allophant.phonemes`:
//   * levensthein_operations / levensthein_matrix (src/edit_distance.rs:116-280 with uniform_costs 483-496): the full f32
//     cost matrix and the FIRST best path with the Rust tie-breaking (deletion only if strictly cheaper than insertion;
//     the diagonal when it is <= that; a diagonal step of equal cost is a match and is not recorded; the walk stops as
//     soon as the remaining cost is 0),
//   * IpaSegmenter (src/ipa_segmenter.rs): leftmost-longest, non-overlapping matches of a segment vocabulary — the
//     semantics of aho_corasick's MatchKind::LeftmostLongest (pinned 0.7 in Cargo.toml) — over UTF-8 bytes, here as a
//     byte trie walked from every unmatched position.
// ALL POINTERS ARE HOST POINTERS.  No CUDA in this file.
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "aph_common.cuh"

namespace aph {

static void edit_matrix(const int64_t* a, int64_t m, const int64_t* b, int64_t n, float* cost) {
  const int64_t w = n + 1;
  for (int64_t j = 0; j <= n; ++j) cost[j] = static_cast<float>(j);
  for (int64_t i = 1; i <= m; ++i) {
    const float* up = cost + (i - 1) * w;
    float* row = cost + i * w;
    row[0] = up[0] + 1.0f;
    for (int64_t j = 1; j <= n; ++j) {
      const float deletion = up[j] + 1.0f;
      const float insertion = row[j - 1] + 1.0f;
      const float substitution = up[j - 1] + (a[i - 1] != b[j - 1] ? 1.0f : 0.0f);
      row[j] = std::min(std::min(insertion, deletion), substitution);
    }
  }
}

struct SegmentTrie {
  struct Node {
    std::vector<std::pair<uint8_t, int32_t>> next;
    bool terminal = false;
  };
  std::vector<Node> nodes;
  SegmentTrie() : nodes(1) {}

  void insert(const uint8_t* bytes, int64_t len) {
    if (len == 0) return;
    int32_t node = 0;
    for (int64_t i = 0; i < len; ++i) {
      int32_t child = -1;
      for (const auto& edge : nodes[node].next)
        if (edge.first == bytes[i]) child = edge.second;
      if (child < 0) {
        child = static_cast<int32_t>(nodes.size());
        nodes[node].next.emplace_back(bytes[i], child);
        nodes.emplace_back();
      }
      node = child;
    }
    nodes[node].terminal = true;
  }

  // end of the longest vocabulary entry that starts at `start`, or -1
  int64_t longest(const uint8_t* text, int64_t len, int64_t start) const {
    int32_t node = 0;
    int64_t best = -1;
    for (int64_t i = start; i < len; ++i) {
      int32_t child = -1;
      for (const auto& edge : nodes[node].next)
        if (edge.first == text[i]) child = edge.second;
      if (child < 0) break;
      node = child;
      if (nodes[node].terminal) best = i + 1;
    }
    return best;
  }
};

}  // namespace aph

using namespace aph;

// (m+1) x (n+1) fp32 cost matrix of uniform-cost Levenshtein (levensthein_matrix, edit_distance.rs:261-269)
extern "C" int aph_edit_matrix(const int64_t* a, int64_t m, const int64_t* b, int64_t n, float* matrix_out) {
  APH_REQUIRE(m >= 0 && n >= 0 && matrix_out && (m == 0 || a) && (n == 0 || b), "edit_matrix: bad arguments");
  edit_matrix(a, m, b, n, matrix_out);
  return APH_OK;
}

// First best path (levensthein_operations, edit_distance.rs:271-280): ops_out receives (action, i, j) triples in FORWARD
// order with action 0 = insertion, 1 = deletion, 2 = substitution (Action::from_int) and (i, j) the matrix coordinates
// AFTER the step; at most m + n of them.  Returns the number of operations (>= 0) or a negative APH_ERR_*.
extern "C" int64_t aph_edit_operations(const int64_t* a, int64_t m, const int64_t* b, int64_t n, int64_t* ops_out, float* final_cost) {
  APH_REQUIRE(m >= 0 && n >= 0 && ops_out && final_cost && (m == 0 || a) && (n == 0 || b), "edit_operations: bad arguments");
  const int64_t w = n + 1;
  std::vector<float> cost(static_cast<size_t>((m + 1) * w));
  edit_matrix(a, m, b, n, cost.data());
  int64_t i = m, j = n, count = 0;
  float current = cost[m * w + n];
  *final_cost = current;
  while (current != 0.0f) {
    int action;  // -1: match
    if (i == 0) {
      if (j == 0) break;
      action = 0;
      current = cost[j - 1];
    } else if (j == 0) {
      action = 1;
      current = cost[(i - 1) * w];
    } else {
      const float deletion = cost[(i - 1) * w + j];
      const float insertion = cost[i * w + j - 1];
      const float diagonal = cost[(i - 1) * w + j - 1];
      float step;
      if (deletion < insertion) {
        action = 1;
        step = deletion;
      } else {
        action = 0;
        step = insertion;
      }
      if (diagonal <= step) {
        action = diagonal == current ? -1 : 2;
        step = diagonal;
      }
      current = step;
    }
    if (action == 0) {
      --j;
    } else if (action == 1) {
      --i;
    } else {
      --i;
      --j;
    }
    if (action >= 0) {
      ops_out[3 * count + 0] = action;
      ops_out[3 * count + 1] = i;
      ops_out[3 * count + 2] = j;
      ++count;
    }
  }
  for (int64_t lo = 0, hi = count - 1; lo < hi; ++lo, --hi)
    for (int c = 0; c < 3; ++c) std::swap(ops_out[3 * lo + c], ops_out[3 * hi + c]);
  return count;
}

// ---- PropertyWeighting (src/edit_distance.rs:497-598): the same three walks over a WEIGHTED matrix.  The substitution cost
// of every (i, j) pair comes in as sub_cost[i * n + j] (the host evaluates property_table[a_i].ne(property_table[b_j]).sum()
// once per distinct symbol pair); row 0 is 0..n in unit steps and column 0 grows by deletion_cost, exactly like
// levensthein_*_general (120-141: `(0..=n).map(|x| x as f32)`, `current_row[0] += deletion_cost`).
static void weighted_matrix(int64_t m, int64_t n, const float* sub_cost, float insertion_cost, float deletion_cost, float* cost) {
  const int64_t w = n + 1;
  for (int64_t j = 0; j <= n; ++j) cost[j] = static_cast<float>(j);
  for (int64_t i = 1; i <= m; ++i) {
    const float* up = cost + (i - 1) * w;
    float* row = cost + i * w;
    row[0] = up[0] + deletion_cost;
    for (int64_t j = 1; j <= n; ++j) {
      const float insertion = row[j - 1] + insertion_cost;
      const float deletion = up[j] + deletion_cost;
      const float substitution = up[j - 1] + sub_cost[(i - 1) * n + (j - 1)];
      row[j] = std::min(std::min(insertion, deletion), substitution);
    }
  }
}

// mode 0: matrix_out gets the (m+1) x (n+1) matrix.  mode 1: ops_out gets the first best path (see aph_edit_operations),
// the return value is their number.  mode 2: stats_out gets (insertions, deletions, substitutions, correct).
extern "C" int64_t aph_edit_weighted(int64_t m, int64_t n, const float* sub_cost, float insertion_cost, float deletion_cost, int32_t mode,
                                     float* matrix_out, int64_t* ops_out, uint64_t* stats_out, float* final_cost) {
  APH_REQUIRE(m >= 0 && n >= 0 && (m == 0 || n == 0 || sub_cost) && mode >= 0 && mode <= 2, "edit_weighted: bad arguments");
  APH_REQUIRE((mode != 0 || matrix_out) && (mode != 1 || ops_out) && (mode != 2 || stats_out), "edit_weighted: missing output");
  const int64_t w = n + 1;
  std::vector<float> local;
  float* cost = matrix_out;
  if (mode != 0) {
    local.resize(static_cast<size_t>((m + 1) * w));
    cost = local.data();
  }
  weighted_matrix(m, n, sub_cost, insertion_cost, deletion_cost, cost);
  if (final_cost) *final_cost = cost[m * w + n];
  if (mode == 0) return 0;
  int64_t i = m, j = n, count = 0;
  uint64_t insertions = 0, deletions = 0, substitutions = 0, correct = 0;
  float current = cost[m * w + n];
  while (current != 0.0f) {
    int action;
    if (i == 0) {
      if (j == 0) break;
      action = 0;
      current = cost[j - 1];
    } else if (j == 0) {
      action = 1;
      current = cost[(i - 1) * w];
    } else {
      const float deletion = cost[(i - 1) * w + j];
      const float insertion = cost[i * w + j - 1];
      const float diagonal = cost[(i - 1) * w + j - 1];
      float step;
      if (deletion < insertion) {
        action = 1;
        step = deletion;
      } else {
        action = 0;
        step = insertion;
      }
      if (diagonal <= step) {
        action = diagonal == current ? -1 : 2;
        step = diagonal;
      }
      current = step;
    }
    if (action == 0) {
      --j;
      ++insertions;
    } else if (action == 1) {
      --i;
      ++deletions;
    } else {
      --i;
      --j;
      if (action == 2) ++substitutions; else ++correct;
    }
    if (mode == 1 && action >= 0) {
      ops_out[3 * count + 0] = action;
      ops_out[3 * count + 1] = i;
      ops_out[3 * count + 2] = j;
      ++count;
    }
  }
  if (mode == 1) {
    for (int64_t lo = 0, hi = count - 1; lo < hi; ++lo, --hi)
      for (int c = 0; c < 3; ++c) std::swap(ops_out[3 * lo + c], ops_out[3 * hi + c]);
    return count;
  }
  stats_out[0] = insertions;
  stats_out[1] = deletions;
  stats_out[2] = substitutions;
  stats_out[3] = correct + static_cast<uint64_t>(i);  // the remaining prefix counts as correct (edit_distance.rs:473-474)
  return 0;
}

// IpaSegmenter::new (ipa_segmenter.rs:96-104): vocabulary as one UTF-8 blob + n+1 byte offsets
extern "C" void* aph_segmenter_create(const char* blob, const int64_t* offsets, int64_t n) {
  if ((n > 0 && (!blob || !offsets)) || n < 0) return nullptr;
  SegmentTrie* trie = new SegmentTrie();
  for (int64_t s = 0; s < n; ++s) trie->insert(reinterpret_cast<const uint8_t*>(blob) + offsets[s], offsets[s + 1] - offsets[s]);
  return trie;
}

extern "C" void aph_segmenter_free(void* handle) { delete static_cast<SegmentTrie*>(handle); }

// find_iter (ipa_segmenter.rs:27-31): byte bounds [start, end) of the leftmost-longest non-overlapping matches, in order.
// bounds_out holds 2 * max_matches entries; returns the number of matches (text_len bounds it) or a negative APH_ERR_*.
extern "C" int64_t aph_segmenter_find(const void* handle, const char* text, int64_t text_len, int64_t* bounds_out, int64_t max_matches) {
  APH_REQUIRE(handle && text_len >= 0 && (text_len == 0 || text) && (max_matches == 0 || bounds_out), "segmenter_find: bad arguments");
  const SegmentTrie* trie = static_cast<const SegmentTrie*>(handle);
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(text);
  int64_t count = 0, position = 0;
  while (position < text_len) {
    const int64_t end = trie->longest(bytes, text_len, position);
    if (end < 0) {
      ++position;
      continue;
    }
    APH_REQUIRE(count < max_matches, "segmenter_find: output buffer too small");
    bounds_out[2 * count] = position;
    bounds_out[2 * count + 1] = end;
    ++count;
    position = end;
  }
  return count;
}
