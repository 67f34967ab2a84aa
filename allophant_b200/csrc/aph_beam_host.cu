// allophant_b200 — host-side CTC beam search: the lexicon-free decoder the reference reaches through
// torchaudio.models.decoder.ctc_decoder(lexicon=None, lm=None, sil_token=blank_token, log_add=True)
// (allophant/predictions.py:210-226).  The algorithm lives in a third-party dependency that is absent here and from
// /root/reference: flashlight-text (LexiconFreeDecoder with ZeroLM; torchaudio >= 2.1 binds flashlight-text 0.0.x).  Its
// published algorithm is restated:
//   * hypotheses are keyed by (token history, last frame's token, "previous frame was blank"); ZeroLM's state IS the token
//     history (a trie node per emitted token), its scores are 0,
//   * per frame every hypothesis is extended by the beam_size_token best tokens; a token starts a new label iff it is not
//     blank and (differs from the previous frame's token or the previous frame was blank); score += emission,
//   * candidates below (best - beam_threshold) are dropped, candidates with equal keys are merged (log-add of the scores,
//     the better one keeps its back pointer), the beam_size best survive,
//   * at the end hypotheses with equal token histories are merged the same way and returned best first.
// The reference passes exp(log-probabilities) as emissions, i.e. path scores are SUMS OF PROBABILITIES; reproduced.
// Parity is UNPINNED (no flashlight here): tests check the decoder against exhaustive enumeration of all alignments.
// Ties between equal scores are broken by std::sort order in flashlight and are not reproducible in general.
// ALL POINTERS ARE HOST POINTERS.  No CUDA in this file.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
#include <unordered_map>
#include <vector>

#include "aph_common.cuh"

namespace aph {

struct BeamHyp {
  double score;
  int32_t history;  // trie node of the emitted token sequence (ZeroLM state)
  int32_t parent;   // index into the previous frame's beam (-1: root)
  int32_t token;    // token of this frame
  bool prev_blank;
};

struct BeamOptions {
  int32_t blank, beam_size, beam_size_token, nbest;
  double beam_threshold;
  bool log_add;
};

struct HistoryTrie {
  std::unordered_map<uint64_t, int32_t> child;
  int32_t nodes = 1;
  int32_t extend(int32_t node, int32_t token) {
    const uint64_t key = (static_cast<uint64_t>(static_cast<uint32_t>(node)) << 32) | static_cast<uint32_t>(token);
    auto found = child.find(key);
    if (found != child.end()) return found->second;
    child.emplace(key, nodes);
    return nodes++;
  }
};

static inline bool same_key(const BeamHyp& a, const BeamHyp& b) { return a.history == b.history && a.token == b.token && a.prev_blank == b.prev_blank; }

// flashlight candidatesStore: threshold, merge equal keys (log-add), keep the `beam` best (sorted best first)
static void store_candidates(std::vector<BeamHyp>& candidates, std::vector<BeamHyp>& out, int32_t beam, double threshold, bool log_add) {
  out.clear();
  std::vector<BeamHyp*> kept;
  kept.reserve(candidates.size());
  for (BeamHyp& c : candidates)
    if (c.score >= threshold) kept.push_back(&c);
  if (kept.empty()) return;
  std::stable_sort(kept.begin(), kept.end(), [](const BeamHyp* a, const BeamHyp* b) {
    if (a->history != b->history) return a->history > b->history;
    if (a->token != b->token) return a->token > b->token;
    if (a->prev_blank != b->prev_blank) return a->prev_blank > b->prev_blank;
    return a->score > b->score;
  });
  size_t merged = 1;
  for (size_t i = 1; i < kept.size(); ++i) {
    if (!same_key(*kept[i], *kept[merged - 1])) {
      kept[merged++] = kept[i];
    } else {
      const double hi = std::max(kept[merged - 1]->score, kept[i]->score);
      const double lo = std::min(kept[merged - 1]->score, kept[i]->score);
      kept[merged - 1]->score = log_add ? hi + std::log1p(std::exp(lo - hi)) : hi;
    }
  }
  kept.resize(merged);
  const size_t final_size = std::min<size_t>(kept.size(), static_cast<size_t>(beam));
  std::partial_sort(kept.begin(), kept.begin() + final_size, kept.end(), [](const BeamHyp* a, const BeamHyp* b) { return a->score > b->score; });
  for (size_t i = 0; i < final_size; ++i) out.push_back(*kept[i]);
}

struct BeamResult {
  std::vector<int64_t> tokens, timesteps;
  double score;
};

static void beam_decode_one(const float* log_emissions, int64_t frames, int32_t classes, const BeamOptions& opt, std::vector<BeamResult>& results) {
  results.clear();
  HistoryTrie trie;
  std::vector<std::vector<BeamHyp>> beams(static_cast<size_t>(frames) + 2);
  beams[0].push_back(BeamHyp{0.0, 0, -1, opt.blank, false});  // decodeBegin: one hypothesis holding the silence (= blank) token
  std::vector<BeamHyp> candidates;
  std::vector<int32_t> order(static_cast<size_t>(classes));
  std::vector<float> emission(static_cast<size_t>(classes));
  const int32_t per_frame = std::min(opt.beam_size_token, classes);
  for (int64_t t = 0; t < frames; ++t) {
    for (int32_t c = 0; c < classes; ++c) {
      emission[c] = std::exp(log_emissions[t * classes + c]);  // predictions.py:226: the decoder is fed probabilities
      order[c] = c;
    }
    if (classes > opt.beam_size_token)
      std::partial_sort(order.begin(), order.begin() + per_frame, order.end(), [&](int32_t l, int32_t r) { return emission[l] > emission[r]; });
    candidates.clear();
    double best = -INFINITY;
    const std::vector<BeamHyp>& previous = beams[t];
    for (size_t p = 0; p < previous.size(); ++p) {
      const BeamHyp& hyp = previous[p];
      for (int32_t r = 0; r < per_frame; ++r) {
        const int32_t n = order[r];
        const double score = hyp.score + static_cast<double>(emission[n]);
        if (score < best - opt.beam_threshold) continue;
        best = std::max(best, score);
        if (n != opt.blank && (n != hyp.token || hyp.prev_blank)) {
          candidates.push_back(BeamHyp{score, trie.extend(hyp.history, n), static_cast<int32_t>(p), n, false});
        } else {
          candidates.push_back(BeamHyp{score, hyp.history, static_cast<int32_t>(p), n, n == opt.blank});
        }
      }
    }
    store_candidates(candidates, beams[t + 1], opt.beam_size, best - opt.beam_threshold, opt.log_add);
  }
  // decodeEnd: every surviving hypothesis gets a final silence frame; equal histories merge
  candidates.clear();
  double best = -INFINITY;
  const std::vector<BeamHyp>& last = beams[frames];
  for (size_t p = 0; p < last.size(); ++p) {
    if (last[p].score < best - opt.beam_threshold) continue;
    best = std::max(best, last[p].score);
    candidates.push_back(BeamHyp{last[p].score, last[p].history, static_cast<int32_t>(p), opt.blank, false});
  }
  store_candidates(candidates, beams[frames + 1], opt.beam_size, best - opt.beam_threshold, opt.log_add);
  const std::vector<BeamHyp>& finals = beams[frames + 1];
  const size_t n_results = std::min<size_t>(finals.size(), static_cast<size_t>(opt.nbest));
  std::vector<int32_t> path(static_cast<size_t>(frames) + 2);
  for (size_t k = 0; k < n_results; ++k) {
    // back pointers give the token of every frame: path[0] = initial silence, path[t + 1] = frame t, path[frames + 1] = final silence
    const BeamHyp* hyp = &finals[k];
    for (int64_t level = frames + 1; level >= 0; --level) {
      path[level] = hyp->token;
      if (level > 0) hyp = &beams[level - 1][hyp->parent];
    }
    BeamResult result;
    result.score = finals[k].score;
    for (int64_t i = 0; i < frames + 2; ++i) {  // torchaudio _get_tokens / _get_timesteps: collapse repeats, drop blanks
      if (path[i] == opt.blank) continue;
      if (i == 0 || path[i] != path[i - 1]) {
        result.tokens.push_back(path[i]);
        result.timesteps.push_back(i);
      }
    }
    results.push_back(std::move(result));
  }
}

}  // namespace aph

using namespace aph;

// log_emissions fp32 [n_seq][t_max][classes] (log-probabilities), lengths int64 [n_seq].  Outputs per sequence and
// rank k < nbest: tokens / timesteps int64 [n_seq][nbest][t_max] (first counts[..] entries valid), counts int64
// [n_seq][nbest] (-1: no such hypothesis), scores fp64 [n_seq][nbest].
extern "C" int aph_ctc_beam_decode(const float* log_emissions, const int64_t* lengths, int64_t n_seq, int64_t t_max, int32_t classes,
                                   int32_t blank, int32_t beam_size, int32_t beam_size_token, double beam_threshold, int32_t nbest,
                                   int32_t log_add, int64_t* tokens_out, int64_t* timesteps_out, int64_t* counts_out, double* scores_out,
                                   int32_t n_threads) {
  APH_REQUIRE(log_emissions && lengths && tokens_out && timesteps_out && counts_out && scores_out, "ctc_beam_decode: null pointer");
  APH_REQUIRE(n_seq >= 0 && t_max >= 0 && classes > 0 && blank >= 0 && blank < classes, "ctc_beam_decode: bad shape");
  APH_REQUIRE(beam_size >= 1 && nbest >= 1 && nbest <= beam_size, "ctc_beam_decode: N-best can not exceed beam width");
  BeamOptions opt{blank, beam_size, beam_size_token > 0 ? beam_size_token : classes, nbest, beam_threshold, log_add != 0};
  for (int64_t s = 0; s < n_seq; ++s) APH_REQUIRE(lengths[s] >= 0 && lengths[s] <= t_max, "ctc_beam_decode: length out of range");
  std::atomic<int64_t> next{0};
  auto worker = [&]() {
    std::vector<BeamResult> results;
    for (;;) {
      const int64_t s = next.fetch_add(1);
      if (s >= n_seq) return;
      beam_decode_one(log_emissions + s * t_max * classes, lengths[s], classes, opt, results);
      for (int32_t k = 0; k < nbest; ++k) {
        const int64_t slot = s * nbest + k;
        if (k >= static_cast<int32_t>(results.size())) {
          counts_out[slot] = -1;
          scores_out[slot] = -INFINITY;
          continue;
        }
        const BeamResult& r = results[k];
        counts_out[slot] = static_cast<int64_t>(r.tokens.size());
        scores_out[slot] = r.score;
        std::copy(r.tokens.begin(), r.tokens.end(), tokens_out + slot * t_max);
        std::copy(r.timesteps.begin(), r.timesteps.end(), timesteps_out + slot * t_max);
      }
    }
  };
  int threads = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  threads = std::max(1, std::min<int>(threads, static_cast<int>(std::max<int64_t>(n_seq, 1))));
  if (threads == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    for (int i = 0; i < threads; ++i) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  return APH_OK;
}
