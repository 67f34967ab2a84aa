/* TEST INFRASTRUCTURE — plain C restatement of the reference's Rust edit distance.
 *
 * Follows /root/reference/src/edit_distance.rs function by function:
 *   ora_levenshtein             levensthein                      70-96   (two usize rows)
 *   ora_levenshtein_statistics  levensthein_statistics_general   372-481 with uniform_costs 483-496, deletion_cost 1 (601-608)
 *   ora_levenshtein_operations  levensthein_operations_general   117-218 (same matrix + backtrace, emits the path)
 *   ora_word_error_rate         EditStatistics::word_error_rate  311-317 (f32 arithmetic)
 * Sequences are int64 symbol ids (the Rust compares Python objects with `!=`).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may link this file; the product's
 * implementation lives in allophant_b200/csrc/aph_edit_distance.cu and is checked AGAINST this one.
 * The Rust source cannot be compiled here (no cargo/rustc in the image), so there is no oracle/_ref.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

uint64_t ora_levenshtein(const int64_t* a, int64_t m, const int64_t* b, int64_t n) {
  uint64_t* previous = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n + 1));
  uint64_t* current = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n + 1));
  for (int64_t j = 0; j <= n; ++j) previous[j] = (uint64_t)j;
  memset(current, 0, sizeof(uint64_t) * (size_t)(n + 1));
  for (int64_t i = 0; i < m; ++i) {
    current[0] = (uint64_t)i + 1;
    for (int64_t j = 0; j < n; ++j) {
      uint64_t deletion = previous[j + 1] + 1;
      uint64_t insertion = current[j] + 1;
      uint64_t substitution = previous[j] + (a[i] != b[j] ? 1u : 0u);
      uint64_t best = deletion < insertion ? deletion : insertion;
      current[j + 1] = best < substitution ? best : substitution;
    }
    uint64_t* swap = previous;
    previous = current;
    current = swap;
  }
  uint64_t result = previous[n];
  free(previous);
  free(current);
  return result;
}

static float* build_matrix(const int64_t* a, int64_t m, const int64_t* b, int64_t n) {
  /* edit_distance.rs:391-412: full (m+1) x (n+1) f32 matrix, row i+1 starts as a clone of row i */
  const int64_t w = n + 1;
  float* matrix = (float*)malloc(sizeof(float) * (size_t)((m + 1) * w));
  for (int64_t j = 0; j <= n; ++j) matrix[j] = (float)j;
  for (int64_t i = 0; i < m; ++i) {
    float* previous = matrix + i * w;
    float* current = matrix + (i + 1) * w;
    memcpy(current, previous, sizeof(float) * (size_t)w);
    current[0] += 1.0f; /* deletion_cost */
    for (int64_t j = 0; j < n; ++j) {
      /* uniform_costs(above = previous[j+1], left = current[j], upper_left = previous[j]) */
      float deletion = previous[j + 1] + 1.0f;
      float insertion = current[j] + 1.0f;
      float substitution = previous[j] + (a[i] != b[j] ? 1.0f : 0.0f);
      float best = insertion < deletion ? insertion : deletion; /* insertion.min(deletion) */
      current[j + 1] = best < substitution ? best : substitution;
    }
  }
  return matrix;
}

/* out = {insertions, deletions, substitutions, correct}; ops (optional, capacity m+n) receives
 * (action, i, j) triples in forward order with action 1 = insertion, 2 = deletion, 3 = substitution. */
static float backtrace(const float* matrix, int64_t m, int64_t n, uint64_t out[4], int64_t* ops, int64_t* n_ops) {
  const int64_t w = n + 1;
  const float final_cost = matrix[m * w + n];
  float current_cost = final_cost;
  int64_t ci = m, cj = n;
  uint64_t insertions = 0, deletions = 0, substitutions = 0, correct = 0;
  int64_t count = 0;
  while (current_cost != 0.0f) {
    int operation; /* 0 = none (match), 1 = insertion, 2 = deletion, 3 = substitution */
    float cost;
    if (ci == 0) {
      if (cj == 0) break;
      operation = 1;
      cost = matrix[ci * w + cj - 1];
    } else if (cj == 0) {
      operation = 2;
      cost = matrix[(ci - 1) * w + cj];
    } else {
      float deletion = matrix[(ci - 1) * w + cj];
      float insertion = matrix[ci * w + cj - 1];
      float substitution = matrix[(ci - 1) * w + cj - 1];
      if (deletion < insertion) {
        operation = 2;
        cost = deletion;
      } else {
        operation = 1;
        cost = insertion;
      }
      if (substitution <= cost) {
        operation = substitution == current_cost ? 0 : 3;
        cost = substitution;
      }
    }
    current_cost = cost;
    switch (operation) {
      case 0: --ci; --cj; ++correct; break;
      case 2: --ci; ++deletions; break;
      case 1: --cj; ++insertions; break;
      default: --ci; --cj; ++substitutions; break;
    }
    if (operation != 0 && ops) {
      ops[3 * count + 0] = operation;
      ops[3 * count + 1] = ci;
      ops[3 * count + 2] = cj;
      ++count;
    }
  }
  correct += (uint64_t)ci; /* edit_distance.rs:473-474 */
  if (out) {
    out[0] = insertions;
    out[1] = deletions;
    out[2] = substitutions;
    out[3] = correct;
  }
  if (ops) {
    for (int64_t lo = 0, hi = count - 1; lo < hi; ++lo, --hi) { /* best_path.reverse() */
      for (int k = 0; k < 3; ++k) {
        int64_t t = ops[3 * lo + k];
        ops[3 * lo + k] = ops[3 * hi + k];
        ops[3 * hi + k] = t;
      }
    }
    *n_ops = count;
  }
  return final_cost;
}

void ora_levenshtein_statistics(const int64_t* a, int64_t m, const int64_t* b, int64_t n, uint64_t out[4]) {
  float* matrix = build_matrix(a, m, b, n);
  backtrace(matrix, m, n, out, NULL, NULL);
  free(matrix);
}

float ora_levenshtein_operations(const int64_t* a, int64_t m, const int64_t* b, int64_t n, int64_t* ops, int64_t* n_ops) {
  float* matrix = build_matrix(a, m, b, n);
  float cost = backtrace(matrix, m, n, NULL, ops, n_ops);
  free(matrix);
  return cost;
}

float ora_word_error_rate(uint64_t insertions, uint64_t deletions, uint64_t substitutions, uint64_t correct) {
  float substituted_or_deleted = (float)(substitutions + deletions);
  return (substituted_or_deleted + (float)insertions) / (substituted_or_deleted + (float)correct);
}
