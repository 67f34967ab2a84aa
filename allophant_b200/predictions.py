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


_PINNED: Dict[Any, List[Tensor]] = {}


def _pinned_like(tensor: Tensor, slot: int) -> Tensor:
    """Rotating pinned host buffers per (shape, dtype): ``cudaHostAlloc`` costs milliseconds, so they are kept."""
    key = (tuple(tensor.shape), tensor.dtype)
    ring = _PINNED.setdefault(key, [])
    while len(ring) <= slot:
        with torch.inference_mode(False):  # the buffers outlive this call: they must not become inference tensors
            ring.append(torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True))
    return ring[slot]


class PendingDecode:
    """Greedy decoding in flight: the collapse kernel and the device-to-host copies are enqueued, nothing has
    synchronised yet.  ``result()`` waits for the copies and builds the hypotheses; a streaming caller launches
    the next batch first, so this host work overlaps the GPU work of that batch."""

    _turn = 0

    def __init__(self, tokens: Tensor, timesteps: Tensor, counts: Tensor, scores: Tensor, n_utt: int, selected: Optional[Dict[str, int]]):
        slot = PendingDecode._turn % 3
        PendingDecode._turn += 1
        # dense packing on the device: the host splits one small tensor instead of mask-indexing n_seq x T elements
        _, packed_tokens, packed_timesteps = ops.ctc_pack_hypotheses(tokens, timesteps, counts)
        self._host = []
        for index, tensor in enumerate((packed_tokens, packed_timesteps, counts, scores)):
            host = _pinned_like(tensor, 4 * slot + index)
            host.copy_(tensor, non_blocking=True)
            self._host.append(host)
        self._event = torch.cuda.Event()
        self._event.record()
        self._n_utt = n_utt
        self._selected = selected
        self._result: Any = None

    def result(self):
        if self._result is None:
            self._event.synchronize()
            per_head = _hypotheses(*self._host, self._n_utt)
            self._result = per_head if self._selected is None else {name: per_head[index] for name, index in self._selected.items()}
        return self._result


def _hypotheses(packed_tokens: Tensor, packed_timesteps: Tensor, counts: Tensor, scores: Tensor, n_utt: int) -> List[List[List[CTCHypothesis]]]:
    """Splits the packed results (already on the host) into per-(head, utterance) hypotheses: two ``Tensor.split``
    calls (C++ loops); the per-hypothesis Python work is only the NamedTuple construction."""
    lengths = counts.tolist()
    total = sum(lengths)
    token_parts = packed_tokens[:total].long().split(lengths)
    timestep_parts = packed_timesteps[:total].long().split(lengths)
    score_parts = scores.clone().unbind(0)
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
        return PendingDecode(tokens, timesteps, counts, scores, n_utt, None).result()[0]


def decode_predictions_async(predictions: Any, names: Optional[Iterable[str]] = None, blank_index: int = 0) -> PendingDecode:
    """Enqueues the greedy decoding of all (or the named) heads of ``Estimator.predict``'s result (one collapse
    launch + four device-to-host copies into pinned buffers) and returns without synchronising."""
    cache = getattr(predictions, "_decode_cache", None)
    if cache is None:
        raise ValueError("decode_predictions_async needs the result of Estimator.predict / Allophant.predict_log_probabilities")
    selected = list(predictions.outputs if names is None else names)
    n_utt, seq = cache["n_utt"], cache["seq"]
    n_heads = cache["argmax"].shape[0]
    tokens, timesteps, counts, scores = ops.ctc_greedy_collapse(
        cache["argmax"], cache["maxlp"], cache["frames32"], n_utt, seq, n_heads * n_utt, blank_index
    )
    return PendingDecode(tokens, timesteps, counts, scores, n_utt, {name: cache["head_index"][name] for name in selected})


def decode_predictions(predictions: Any, names: Optional[Iterable[str]] = None, blank_index: int = 0) -> Dict[str, List[List[CTCHypothesis]]]:
    """Greedy-decodes all (or the named) heads of ``Estimator.predict``'s result in one launch.

    Equivalent to ``{name: GreedyCTCDecoder()(predictions.outputs[name].transpose(1, 0).contiguous(),
    predictions.lengths) for name in names}`` (``run.py:767-774``)."""
    if getattr(predictions, "_decode_cache", None) is None:
        decoder = GreedyCTCDecoder(blank_index)
        selected = list(predictions.outputs if names is None else names)
        return {name: decoder(predictions.outputs[name].transpose(1, 0), predictions.lengths) for name in selected}
    return decode_predictions_async(predictions, names, blank_index).result()


class BeamCTCDecoder:
    """CTC beam search (``predictions.py:210-226``): the lexicon-free flashlight decoder behind
    ``torchaudio.models.decoder.ctc_decoder(lexicon=None, lm=None, sil_token=blank, beam_size=beam_width, nbest=n_best,
    log_add=True)`` with torchaudio's defaults (``beam_size_token`` = all tokens, ``beam_threshold`` = 50), restated in host
    C++ (``aph_ctc_beam_decode``) and run over (utterance) threads.  Like the reference it scores paths by the SUM OF
    PROBABILITIES (the decoder is fed ``log_emissions.exp()``)."""

    BEAM_THRESHOLD = 50.0

    def __init__(self, tokens: List[str], beam_width: int, n_best: int = 1, blank_index: int = 0) -> None:
        self._tokens = list(tokens)
        self._beam_width, self._n_best, self._blank_index = beam_width, n_best, blank_index

    def __call__(self, log_emissions: Tensor, lengths: Optional[Tensor] = None) -> List[List[CTCHypothesis]]:
        import ctypes

        from ._lib import check, lib

        emissions = log_emissions.detach().to(device="cpu", dtype=torch.float32).contiguous()
        n_seq, t_max, classes = emissions.shape
        if lengths is None:
            lengths = torch.full((n_seq,), t_max, dtype=torch.int64)
        host_lengths = lengths.detach().to(device="cpu", dtype=torch.int64).contiguous()
        tokens = torch.zeros(n_seq, self._n_best, max(1, t_max), dtype=torch.int64)
        timesteps = torch.zeros_like(tokens)
        counts = torch.zeros(n_seq, self._n_best, dtype=torch.int64)
        scores = torch.zeros(n_seq, self._n_best, dtype=torch.float64)
        pointer = lambda tensor: ctypes.c_void_p(tensor.data_ptr())  # noqa: E731
        check(
            lib.aph_ctc_beam_decode(
                pointer(emissions), pointer(host_lengths), n_seq, t_max, classes, self._blank_index, self._beam_width, 0, self.BEAM_THRESHOLD,
                self._n_best, 1, pointer(tokens), pointer(timesteps), pointer(counts), pointer(scores), 0,
            ),  # fmt: skip
            "aph_ctc_beam_decode",
        )
        results: List[List[CTCHypothesis]] = []
        for sequence in range(n_seq):
            hypotheses = []
            for rank in range(self._n_best):
                count = int(counts[sequence, rank])
                if count < 0:
                    break
                hypotheses.append(
                    CTCHypothesis(tokens[sequence, rank, :count].clone(), [], float(scores[sequence, rank]), timesteps[sequence, rank, :count].to(torch.int32))
                )
            results.append(hypotheses)
        return results


def _ctc_decoder(categories: Iterable[str], beam_width: int = 1, n_best: int = 1) -> Any:
    assert n_best <= beam_width, "N-best can not exceed beam width"
    # Optimized decoder for greedy decoding with log probabilities
    if beam_width == 1:
        return GreedyCTCDecoder()
    return BeamCTCDecoder(["<blank>", *categories], beam_width, n_best)


def feature_decoders(indexer: Any, beam_width: int = 1, feature_names: Optional[Iterable[str]] = None, n_best: int = 1) -> Dict[str, GreedyCTCDecoder]:
    return {
        name: _ctc_decoder(indexer.feature_categories(name), beam_width, n_best)
        for name in (indexer.feature_names if feature_names is None else feature_names)
    }


# ----------------------------------------------------------------------------------------------------------------------
# Prediction files: JSON lines, format 1.1.0 (allophant/predictions.py:29-186).  First line = metadata, then one
# UtterancePrediction (or UtteranceEdits) per line; optional gzip, inferred from a ".gz" suffix.
# ----------------------------------------------------------------------------------------------------------------------
import dataclasses as _dataclasses
import gzip as _gzip
import io as _io
import json as _json
import os as _os
from dataclasses import dataclass as _dataclass
from typing import Iterator, Tuple, Union

from . import __version__ as _PACKAGE_VERSION
from . import phonemes as _phonemes
from .config import FeatureSet

CURRENT_FORMAT_VERSION = (1, 1, 0)
SUPPORTED_VERSIONS = [CURRENT_FORMAT_VERSION]


@_dataclass
class PredictionMetaData:
    """``predictions.py:34-47``; ``indexer_state`` is the serialised state of the attribute indexer (a plain dict here)."""

    prediction_arguments: str
    corpus_type: str
    languages: List[str]
    feature_set: FeatureSet
    indexer_state: Any
    classifiers: List[str]
    label_inventories: Optional[Dict[str, List[str]]] = None
    package_version: str = _PACKAGE_VERSION
    format_version: Tuple[int, int, int] = CURRENT_FORMAT_VERSION

    def dumps(self) -> str:
        fields = _dataclasses.asdict(self)
        fields["feature_set"] = self.feature_set.value
        fields["format_version"] = list(self.format_version)
        return _json.dumps(fields)

    @classmethod
    def loads(cls, line: str) -> "PredictionMetaData":
        fields = _json.loads(line)
        fields["feature_set"] = FeatureSet(fields["feature_set"])
        fields["format_version"] = tuple(fields.get("format_version", CURRENT_FORMAT_VERSION))
        if fields["format_version"] not in SUPPORTED_VERSIONS:
            raise ValueError(f"Unsupported prediction format version {fields['format_version']}, supported: {SUPPORTED_VERSIONS}")
        return cls(**fields)


@_dataclass
class UtterancePrediction:
    """``predictions.py:50-55``: ``predictions`` maps a classifier name to its n-best label sequences."""

    language: str
    utterance_id: str
    predictions: Dict[str, List[List[str]]]
    labels: Optional[List[List[str]]] = None

    def to_json(self) -> str:
        return _json.dumps(_dataclasses.asdict(self))

    @classmethod
    def from_json(cls, line: str) -> "UtterancePrediction":
        return cls(**_json.loads(line))


@_dataclass
class UtteranceEdits:
    """``predictions.py:68-79``: edit operations per classifier; actions are stored as integers (``int(Action)``)."""

    language: str
    utterance_id: str
    expected: Dict[str, List[str]]
    edit_operations: Dict[str, List[Tuple[Any, str, str]]]

    def to_json(self) -> str:
        fields = _dataclasses.asdict(self)
        fields["edit_operations"] = {
            name: [[int(action), expected, actual] for action, expected, actual in operations] for name, operations in self.edit_operations.items()
        }
        return _json.dumps(fields)

    @classmethod
    def from_json(cls, line: str) -> "UtteranceEdits":
        fields = _json.loads(line)
        fields["edit_operations"] = {
            name: [(_phonemes.Action.from_int(action), expected, actual) for action, expected, actual in operations]
            for name, operations in fields["edit_operations"].items()
        }
        return cls(**fields)


def levensthein_substitutions(expected: List[str], actual: List[str]) -> List[Tuple[Any, str, str]]:
    """``predictions.py:58-59``."""
    return _phonemes.to_substitutions(expected, actual, _phonemes.levensthein_operations(expected, actual)[0])


PathOrFileBinary = Union[str, "_os.PathLike[str]", Any]


def _file_path(file: PathOrFileBinary) -> str:
    if isinstance(file, (str, _os.PathLike)):
        return _os.fspath(file)
    return getattr(file, "name", "")


def _infer_gzip(file: PathOrFileBinary) -> bool:
    return _os.path.splitext(_file_path(file))[1] == ".gz"


def _binary(file: PathOrFileBinary, mode: str):
    return open(file, mode) if isinstance(file, (str, _os.PathLike)) else file


class JsonlWriter:
    """``predictions.py:159-186``: exclusive creation ("x"), metadata line first."""

    def __init__(self, file: PathOrFileBinary, metadata: PredictionMetaData, gzip: Optional[bool] = False) -> None:
        self._wrapped_file = file
        self._gzip = _infer_gzip(file) if gzip is None else gzip
        self._meta_data = metadata

    def __enter__(self) -> "JsonlWriter":
        raw = _gzip.open(self._wrapped_file, "x") if self._gzip else _binary(self._wrapped_file, "xb")
        self._file = _io.TextIOWrapper(raw, encoding="utf-8")
        self._file.write(self._meta_data.dumps() + "\n")
        return self

    def __exit__(self, *_: Any) -> None:
        self._file.close()

    def write(self, serialized: Any) -> None:
        self._file.write(str(serialized.to_json()) + "\n")


class JsonlReader:
    """``predictions.py:95-133``."""

    def __init__(self, file: PathOrFileBinary, gzip: Optional[bool] = None) -> None:
        self._wrapped_file = file
        self._gzip = _infer_gzip(file) if gzip is None else gzip

    def read_meta(self) -> Any:
        return None

    def process_line(self, line: str) -> Any:
        return line

    def __iter__(self) -> Iterator[Any]:
        for line in self._file:
            yield self.process_line(line)

    def __enter__(self):
        raw = _gzip.open(self._wrapped_file, "r") if self._gzip else _binary(self._wrapped_file, "rb")
        self._file = _io.TextIOWrapper(raw, encoding="utf-8")
        self._metadata = self.read_meta()
        return self

    def __exit__(self, *_: Any) -> None:
        self._file.close()


class PredictionReader(JsonlReader):
    def read_meta(self) -> PredictionMetaData:
        return PredictionMetaData.loads(self._file.readline())

    @property
    def metadata(self) -> PredictionMetaData:
        return self._metadata

    def process_line(self, line: str) -> UtterancePrediction:
        return UtterancePrediction.from_json(line)


class StatisticsReader(JsonlReader):
    def read_meta(self) -> PredictionMetaData:
        return PredictionMetaData.loads(self._file.readline())

    @property
    def metadata(self) -> PredictionMetaData:
        return self._metadata

    def process_line(self, line: str) -> UtteranceEdits:
        return UtteranceEdits.from_json(line)
