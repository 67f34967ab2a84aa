"""Greedy CTC decoding (drop-in for ``allophant/predictions.py:189-254``).

``GreedyCTCDecoder()(log_emissions, lengths)`` keeps the reference's per-head call signature
and result type; ``decode_predictions`` decodes every head of a ``Predictions`` object with one
kernel launch and one device-to-host copy instead of 37 x N Python iterations.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, NamedTuple, Optional

import torch
from torch import Tensor

from . import ops


class CTCHypothesis(NamedTuple):
    """Same fields as ``torchaudio.models.decoder.CTCHypothesis`` (which cannot be imported without
    flashlight-text): tokens, words, score, 1-based timesteps of the kept run starts."""

    tokens: torch.LongTensor
    words: List[str]
    score: float
    timesteps: torch.IntTensor


def _hypotheses(tokens: Tensor, timesteps: Tensor, counts: Tensor, scores: Tensor, n_utt: int) -> List[List[List[CTCHypothesis]]]:
    """Splits the padded device results into per-(head, utterance) hypotheses on the host.

    One device-to-host copy per array, then a single masked compaction and ``Tensor.split`` (C++ loops):
    the per-hypothesis Python work is only the NamedTuple construction."""
    counts_h = counts.cpu()
    lengths = counts_h.tolist()
    frames = tokens.shape[1]
    keep = torch.arange(frames).unsqueeze(0) < counts_h.unsqueeze(1)
    token_parts = tokens.cpu()[keep].long().split(lengths)
    timestep_parts = timesteps.cpu()[keep].long().split(lengths)
    score_parts = scores.cpu().unbind(0)
    n_heads = len(lengths) // n_utt
    result = []
    for head in range(n_heads):
        base = head * n_utt
        result.append(
            [[CTCHypothesis(token_parts[base + utt], [], score_parts[base + utt], timestep_parts[base + utt])] for utt in range(n_utt)]
        )
    return result


class GreedyCTCDecoder:
    def __init__(self, blank_index: int = 0):
        super().__init__()
        self._blank_index = blank_index

    def __call__(self, log_emissions: Tensor, lengths: Tensor) -> List[List[CTCHypothesis]]:
        """``log_emissions`` fp32 ``[N, T', classes]`` on the GPU, ``lengths`` ``[N]`` frame counts."""
        if not log_emissions.is_cuda:
            raise RuntimeError("allophant_b200.GreedyCTCDecoder decodes on the GPU only (no CPU fallback exists)")
        if log_emissions.dim() != 3:
            raise ValueError(f"expected log emissions of shape [batch, frames, classes], got {tuple(log_emissions.shape)}")
        emissions = log_emissions.float()
        if emissions.stride(-1) != 1 or emissions.stride(1) != emissions.shape[2] or emissions.stride(0) != emissions.shape[1] * emissions.shape[2]:
            emissions = emissions.contiguous()
        n_utt, seq, classes = emissions.shape
        device = emissions.device
        rows = n_utt * seq
        argmax = torch.empty(rows, device=device, dtype=torch.int32)
        maxlp = torch.empty(rows, device=device, dtype=torch.float32)
        ops.argmax_rows(emissions, classes, rows, classes, argmax, maxlp)
        frames32 = lengths.to(device=device, dtype=torch.int32).contiguous()
        tokens, timesteps, counts, scores = ops.ctc_greedy_collapse(argmax, maxlp, frames32, n_utt, seq, n_utt, self._blank_index)
        return _hypotheses(tokens, timesteps, counts, scores, n_utt)[0]


def decode_predictions(predictions: Any, names: Optional[Iterable[str]] = None, blank_index: int = 0) -> Dict[str, List[List[CTCHypothesis]]]:
    """Greedy-decodes all (or the named) heads of ``Estimator.predict``'s result in one launch.

    Equivalent to ``{name: GreedyCTCDecoder()(predictions.outputs[name].transpose(1, 0).contiguous(),
    predictions.lengths) for name in names}`` (``run.py:767-774``)."""
    cache = getattr(predictions, "_decode_cache", None)
    selected = list(predictions.outputs if names is None else names)
    if cache is None:
        decoder = GreedyCTCDecoder(blank_index)
        return {name: decoder(predictions.outputs[name].transpose(1, 0), predictions.lengths) for name in selected}
    n_utt, seq = cache["n_utt"], cache["seq"]
    n_heads = cache["argmax"].shape[0]
    tokens, timesteps, counts, scores = ops.ctc_greedy_collapse(
        cache["argmax"], cache["maxlp"], cache["frames32"], n_utt, seq, n_heads * n_utt, blank_index
    )
    per_head = _hypotheses(tokens, timesteps, counts, scores, n_utt)
    return {name: per_head[cache["head_index"][name]] for name in selected}


def _ctc_decoder(categories: Iterable[str], beam_width: int = 1, n_best: int = 1) -> GreedyCTCDecoder:
    assert n_best <= beam_width, "N-best can not exceed beam width"
    if beam_width == 1:
        return GreedyCTCDecoder()
    raise NotImplementedError("beam search decoding (flashlight-text) is outside this build's hot path; use beam_width=1")


def feature_decoders(indexer: Any, beam_width: int = 1, feature_names: Optional[Iterable[str]] = None, n_best: int = 1) -> Dict[str, GreedyCTCDecoder]:
    return {
        name: _ctc_decoder(indexer.feature_categories(name), beam_width, n_best)
        for name in (indexer.feature_names if feature_names is None else feature_names)
    }
