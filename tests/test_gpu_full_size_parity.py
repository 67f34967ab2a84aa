"""-m gpu: BASELINE.json's full-size configurations VALUE-compared with the fp32 CPU oracle (the restatement that is bit-identical
to the unmodified reference on every golden case), on identical seeded weights and inputs:

* configs[1]: Multitask, XLS-R-300M shape (24 layers), 32 x 10 s ragged, 37 heads;
* a configs[3] slice: Hierarchical phoneme head (OUTPUT + 36 attribute posteriors) composed over an inventory of V = 3 183
  phones, 4 x 10 s;
* configs[4]: a 30 s utterance (T' = 1 499) next to a shorter one.

Per head the report holds: the range error (max |ours - oracle| / max |oracle| over the valid frames), the largest
probability-space error, the mean per-frame KL divergence, the number of frames whose argmax differs from the oracle's
(and the oracle's top-2 margin on those frames), whether the greedy hypotheses are identical, and PER / AER on BOTH sides
(``estimator.py:1035-1046`` -> ``predictions.py:194-207`` -> ``src/edit_distance.rs:601-608``) against the same synthetic
transcripts.  The bounds asserted below are the stated tolerances of this path (bf16 GEMM operands, fp32 accumulation, fp32
residual stream and statistics):

    range error                     < 2e-2      (north_star: 2e-2 for bf16 kernels; measured <= 1.2e-2)
    probability-space error         < 1.5e-2    absolute, any class of any valid frame, attribute heads (measured 7e-3);
                                    < 5e-2      composed phoneme head (two chained bf16 contractions; measured 3.2e-2)
    mean KL(oracle || ours)         < 1e-3      nats per frame (measured 1.9e-4)
    argmax flips                    <= 2 %      of the valid frames of an attribute head (measured 1.2 % worst, 0.64 % over all heads),
                                    <= 4 %      of the phoneme head's (measured 2.2 % over the 3 184-phone inventory, 0.6 % over 26), every
                                                one on an oracle top-2 margin < 0.05 nats (attribute heads) / 0.15 nats (phoneme)
    PER / AER difference            <= 5e-2     absolute, per head, ours vs oracle against the same transcripts (measured 3.0e-2)
    hypothesis disagreement         <= 1e-1     edit distance ours vs oracle / oracle hypothesis length, per head (measured 6.1e-2)

Why flips exist at all: the models are RANDOM-INIT (no checkpoint can be fetched), so every head's posterior is nearly flat and
~1 % of the frames have a top-2 margin below the bf16 noise of a 24-layer encoder (log-probabilities agree to ~1e-2 of their
range = 0.02-0.03 nats).  Given the SAME log-probabilities the decoder is bit-exact (tests/test_gpu_e2e_parity.py).
The report is written to ``gpurun_out/r02_full_size_parity.json`` (copied to ``profiles/`` by the round script).
"""
import json
import os
import time

import pytest
import torch

from oracle import restatement
from tests import helpers

pytestmark = pytest.mark.gpu

RANGE_TOL = 2e-2
PROB_TOL = {"attribute": 1.5e-2, "phoneme": 5e-2}
KL_TOL = 1e-3
FLIP_RATE_TOL = {"attribute": 2e-2, "phoneme": 4e-2}
FLIP_MARGIN_TOL = {"attribute": 0.05, "phoneme": 0.15}
ERROR_RATE_TOL = 5e-2
DISAGREEMENT_TOL = 1e-1

REPORT_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_full_size_parity.json")
_REPORT = {}


def _write_report() -> None:
    directory = os.path.dirname(REPORT_PATH)
    if os.path.isdir(directory):
        with open(REPORT_PATH, "w") as file:
            json.dump(_REPORT, file, indent=1, sort_keys=True)


def compare_with_oracle(case: str, spec, audio, lengths, tfi=None):
    """Runs the oracle on the host and the CUDA path on the device and returns the per-head report."""
    from allophant_b200 import phonemes
    from allophant_b200.dataset_processing import Batch
    from allophant_b200.predictions import decode_predictions

    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    started = time.perf_counter()
    reference, frames = oracle.predict(audio, lengths, None, tfi)
    oracle_seconds = time.perf_counter() - started
    batch = Batch(audio.cuda(), lengths.cuda(), torch.zeros(len(lengths), dtype=torch.long).cuda())
    with torch.inference_mode():
        predictions = model.predict_log_probabilities(batch, None if tfi is None else tfi.cuda())
        decoded = decode_predictions(predictions)
    assert torch.equal(predictions.lengths.cpu(), frames)
    assert list(predictions.outputs) == list(reference)
    counts = frames.tolist()
    heads = {}
    for index, (name, expected) in enumerate(reference.items()):
        ours = predictions.outputs[name].float().cpu()
        assert ours.shape == expected.shape, (name, ours.shape, expected.shape)
        classes = expected.shape[-1]
        worst = scale = prob = kl = 0.0
        flips = valid = 0
        margin = 0.0
        for utterance, count in enumerate(counts):
            a, b = ours[:count, utterance], expected[:count, utterance]
            worst = max(worst, float((a - b).abs().max()))
            scale = max(scale, float(b.abs().max()))
            prob = max(prob, float((a.exp() - b.exp()).abs().max()))
            kl += float((b.exp() * (b - a)).sum())
            differs = a.argmax(-1) != b.argmax(-1)
            valid += count
            flips += int(differs.sum())
            if differs.any():
                top2 = b.topk(2, -1).values
                margin = max(margin, float((top2[:, 0] - top2[:, 1])[differs].max()))
        # greedy hypotheses of both sides (predictions.py:194-207) and PER / AER against the same synthetic transcripts
        theirs = restatement.greedy_ctc_decode(expected.transpose(0, 1), frames)
        labels, label_lengths = restatement.synthetic_labels(frames, classes, seed=500 + index)
        transcripts = [labels[row, : int(label_lengths[row])].tolist() for row in range(len(counts))]
        ours_tokens = [hypothesis[0].tokens.tolist() for hypothesis in decoded[name]]
        their_tokens = [hypothesis[0].tokens.tolist() for hypothesis in theirs]
        ours_stats, their_stats, between = phonemes.EditStatistics.zeros(), phonemes.EditStatistics.zeros(), phonemes.EditStatistics.zeros()
        for statistics in phonemes.levensthein_statistics_batch(list(zip(transcripts, ours_tokens))):
            ours_stats += statistics
        for statistics in phonemes.levensthein_statistics_batch(list(zip(transcripts, their_tokens))):
            their_stats += statistics
        for statistics in phonemes.levensthein_statistics_batch(list(zip(their_tokens, ours_tokens))):
            between += statistics
        # timesteps / scores of the identical hypotheses agree as well
        identical = sum(int(a == b) for a, b in zip(ours_tokens, their_tokens))
        heads[name] = dict(
            classes=classes, range_error=worst / scale, probability_error=prob, mean_kl=kl / valid, frames=valid, argmax_flips=flips,
            flip_rate=flips / valid, worst_flip_margin_nats=margin, identical_hypotheses=identical, utterances=len(counts),
            error_rate_ours=ours_stats.word_error_rate(), error_rate_oracle=their_stats.word_error_rate(),
            disagreement=(between.word_error_rate() if sum(map(len, their_tokens)) else 0.0),
            tokens_oracle=sum(map(len, their_tokens)), tokens_ours=sum(map(len, ours_tokens)),
        )  # fmt: skip
    summary = dict(
        heads=len(heads), frames_per_head=sum(counts), oracle_seconds=oracle_seconds,
        worst_range_error=max(h["range_error"] for h in heads.values()),
        worst_probability_error=max(h["probability_error"] for h in heads.values()),
        worst_mean_kl=max(h["mean_kl"] for h in heads.values()),
        total_flips=sum(h["argmax_flips"] for h in heads.values()), total_frames=sum(h["frames"] for h in heads.values()),
        worst_flip_rate=max(h["flip_rate"] for h in heads.values()),
        worst_flip_margin_nats=max(h["worst_flip_margin_nats"] for h in heads.values()),
        worst_error_rate_difference=max(abs(h["error_rate_ours"] - h["error_rate_oracle"]) for h in heads.values()),
        worst_disagreement=max(h["disagreement"] for h in heads.values()),
        identical_hypotheses=sum(h["identical_hypotheses"] for h in heads.values()), hypotheses=len(heads) * len(counts),
        phoneme=heads.get("phoneme"),
    )  # fmt: skip
    _REPORT[case] = dict(summary=summary, heads=heads)
    _write_report()
    print(f"[{case}] " + json.dumps({k: v for k, v in summary.items() if k != "phoneme"}))
    return summary, heads


def check_bounds(summary, heads) -> None:
    for name, head in heads.items():
        kind = "phoneme" if name == "phoneme" else "attribute"
        assert head["range_error"] < RANGE_TOL, (name, head)
        assert head["probability_error"] < PROB_TOL[kind], (name, head)
        assert head["mean_kl"] < KL_TOL, (name, head)
        assert head["flip_rate"] <= FLIP_RATE_TOL[kind], (name, head)
        assert head["worst_flip_margin_nats"] < FLIP_MARGIN_TOL[kind], (name, head)
        assert abs(head["error_rate_ours"] - head["error_rate_oracle"]) <= ERROR_RATE_TOL, (name, head)
        assert head["disagreement"] <= DISAGREEMENT_TOL, (name, head)


def test_config1_multitask_32x10s_matches_the_oracle():
    """BASELINE configs[1]: 24 layers, 32 x 10 s (three utterances shorter: 5 s, 1 s, 0.085 s), 37 heads, inventory of 25."""
    spec = restatement.multitask_spec(n_train_phonemes=60)
    samples = 160_000
    lengths = torch.full((32,), samples, dtype=torch.long)
    lengths[3], lengths[17], lengths[31] = samples // 2, 16000, 400 + 3 * 320
    audio = restatement.synthetic_audio(32, samples, seed=7) * restatement.mask_sequence(lengths)
    tfi = torch.randint(0, 3, (25, 36), generator=torch.Generator().manual_seed(1))
    summary, heads = compare_with_oracle("config1_multitask_32x10s", spec, audio, lengths, tfi)
    assert summary["heads"] == 37 and summary["frames_per_head"] == 29 * 499 + 249 + 49 + 4
    check_bounds(summary, heads)


def test_config3_hierarchical_inventory_3183_matches_the_oracle():
    """A slice of BASELINE configs[3]: the hierarchical phoneme head over V = 3 183 composed phone embeddings, 4 x 10 s."""
    spec = restatement.multitask_spec(n_train_phonemes=60, hierarchical=True)
    samples = 160_000
    lengths = torch.tensor([samples, samples - 12_345, samples, 96_000])
    audio = restatement.synthetic_audio(4, samples, seed=11) * restatement.mask_sequence(lengths)
    tfi = torch.randint(0, 3, (3183, 36), generator=torch.Generator().manual_seed(5))
    summary, heads = compare_with_oracle("config3_hierarchical_v3183_4x10s", spec, audio, lengths, tfi)
    assert heads["phoneme"]["classes"] == 3184
    check_bounds(summary, heads)


def test_config4_30s_utterance_matches_the_oracle():
    """BASELINE configs[4]: 30 s utterances (T' = 1 499: 24 key blocks per query tile) next to a 17 s one."""
    spec = restatement.multitask_spec(n_train_phonemes=60)
    samples = 480_000
    lengths = torch.tensor([samples, 272_000])
    audio = restatement.synthetic_audio(2, samples, seed=13) * restatement.mask_sequence(lengths)
    tfi = torch.randint(0, 3, (25, 36), generator=torch.Generator().manual_seed(1))
    summary, heads = compare_with_oracle("config4_30s", spec, audio, lengths, tfi)
    assert summary["frames_per_head"] == 1499 + 849
    check_bounds(summary, heads)
