"""Drop-in for the edit-distance half of the reference's Rust extension ``allophant.phonemes``
(``src/edit_distance.rs``; stub ``allophant/phonemes.pyi``): ``levensthein``,
``levensthein_statistics``, ``EditStatistics``, plus a batched entry point for whole evaluations.

Symbols can be any hashable Python objects (the Rust compares them with ``!=``); they are mapped
to integer ids and the dynamic programme runs in ``liballophant_b200.so`` on the host.
"""
from __future__ import annotations

import ctypes
from enum import Enum
from typing import Dict, Hashable, Iterable, List, Sequence, Tuple

import numpy as np

from ._lib import check, lib


class Action(Enum):
    """``src/edit_distance.rs:41-69``: the PyO3 enum's integer values are its Rust discriminants 0 / 1 / 2 (the ``.pyi`` stub of
    the reference lists 1 / 2 / 3, but ``from_int`` and ``int()`` of the compiled extension use the discriminants)."""

    INSERTION = 0
    DELETION = 1
    SUBSTITUTION = 2

    @staticmethod
    def from_int(integer: int) -> "Action":
        if integer not in (0, 1, 2):
            raise ValueError(f"Invalid enum value {integer}")
        return Action(integer)

    def __int__(self) -> int:
        return self.value


class EditStatistics:
    __slots__ = ("insertions", "deletions", "substitutions", "correct")

    def __init__(self, insertions: int, deletions: int, substitutions: int, correct: int) -> None:
        self.insertions = int(insertions)
        self.deletions = int(deletions)
        self.substitutions = int(substitutions)
        self.correct = int(correct)

    @classmethod
    def zeros(cls) -> "EditStatistics":
        return cls(0, 0, 0, 0)

    def word_error_rate(self) -> float:
        """(S + D + I) / (S + D + C), evaluated in f32 like the Rust (``edit_distance.rs:311-317``)."""
        return float(lib.aph_word_error_rate(self.insertions, self.deletions, self.substitutions, self.correct))

    def _expected_count(self) -> np.float32:
        return np.float32(self.substitutions + self.deletions + self.correct)

    def substitution_rate(self) -> float:
        return float(np.float32(self.substitutions) / self._expected_count())

    def insertion_rate(self) -> float:
        return float(np.float32(self.insertions) / self._expected_count())

    def deletion_rate(self) -> float:
        return float(np.float32(self.deletions) / self._expected_count())

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, EditStatistics):
            return NotImplemented
        return tuple(self) == tuple(other)

    def __hash__(self) -> int:
        return hash(tuple(self))

    def __iter__(self):
        return iter((self.insertions, self.deletions, self.substitutions, self.correct))

    def __repr__(self) -> str:
        return (
            f"EditStatistics(insertions={self.insertions}, deletions={self.deletions}, "
            f"substitutions={self.substitutions}, correct={self.correct})"
        )

    def __add__(self, rhs: "EditStatistics") -> "EditStatistics":
        return EditStatistics(
            self.insertions + rhs.insertions, self.deletions + rhs.deletions, self.substitutions + rhs.substitutions, self.correct + rhs.correct
        )

    def __iadd__(self, other: "EditStatistics") -> "EditStatistics":
        self.insertions += other.insertions
        self.deletions += other.deletions
        self.substitutions += other.substitutions
        self.correct += other.correct
        return self


def _encode(pairs: Sequence[Tuple[Sequence[Hashable], Sequence[Hashable]]]):
    ids: Dict[Hashable, int] = {}
    expected: List[int] = []
    actual: List[int] = []
    expected_offsets = [0]
    actual_offsets = [0]
    for a, b in pairs:
        for symbol in a:
            expected.append(ids.setdefault(_key(symbol), len(ids)))
        for symbol in b:
            actual.append(ids.setdefault(_key(symbol), len(ids)))
        expected_offsets.append(len(expected))
        actual_offsets.append(len(actual))
    to_array = lambda values: np.ascontiguousarray(np.asarray(values if values else [0], dtype=np.int64))  # noqa: E731
    return to_array(expected), to_array(expected_offsets), to_array(actual), to_array(actual_offsets)


def _key(symbol: Hashable) -> Hashable:
    # tensors / numpy scalars compare by value in the reference (`!=` on the objects)
    item = getattr(symbol, "item", None)
    if callable(item):
        try:
            return item()
        except (ValueError, RuntimeError):
            pass
    return symbol


def _pointer(array: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(array.ctypes.data)


def levensthein_statistics_batch(pairs: Sequence[Tuple[Sequence[Hashable], Sequence[Hashable]]], threads: int = 0) -> List[EditStatistics]:
    """``[levensthein_statistics(expected, actual) for expected, actual in pairs]`` in one native call."""
    if not pairs:
        return []
    expected, expected_offsets, actual, actual_offsets = _encode(pairs)
    out = np.zeros((len(pairs), 4), dtype=np.uint64)
    check(
        lib.aph_edit_statistics_batch(_pointer(expected), _pointer(expected_offsets), _pointer(actual), _pointer(actual_offsets), len(pairs), _pointer(out), None, threads),
        "aph_edit_statistics_batch",
    )
    return [EditStatistics(*row) for row in out.tolist()]


def levensthein_statistics(string_a: Sequence[Hashable], string_b: Sequence[Hashable]) -> EditStatistics:
    return levensthein_statistics_batch([(string_a, string_b)], threads=1)[0]


def levensthein(string_a: Sequence[Hashable], string_b: Sequence[Hashable]) -> int:
    expected, expected_offsets, actual, actual_offsets = _encode([(string_a, string_b)])
    out = np.zeros(1, dtype=np.uint64)
    check(
        lib.aph_edit_statistics_batch(_pointer(expected), _pointer(expected_offsets), _pointer(actual), _pointer(actual_offsets), 1, None, _pointer(out), 1),
        "aph_edit_statistics_batch",
    )
    return int(out[0])


LevenstheinOperations = List[Tuple[Action, int, int]]


def levensthein_operations(string_a: Sequence[Hashable], string_b: Sequence[Hashable]) -> Tuple[LevenstheinOperations, float]:
    """First best edit path and its cost (``src/edit_distance.rs:116-218, 271-280``): ``(action, i, j)`` with the matrix
    coordinates AFTER each step, matches are not listed."""
    expected, _, actual, _ = _encode([(string_a, string_b)])
    m, n = len(string_a), len(string_b)
    out = np.zeros((max(1, m + n), 3), dtype=np.int64)
    cost = ctypes.c_float(0.0)
    count = lib.aph_edit_operations(_pointer(expected), m, _pointer(actual), n, _pointer(out), ctypes.byref(cost))
    if count < 0:
        check(int(count), "aph_edit_operations")
    return [(Action.from_int(int(action)), int(i), int(j)) for action, i, j in out[:count].tolist()], float(cost.value)


def levensthein_matrix(string_a: Sequence[Hashable], string_b: Sequence[Hashable]):
    """The full fp32 cost matrix as a ``torch.Tensor`` ``[len(a) + 1, len(b) + 1]`` (``src/edit_distance.rs:220-269``)."""
    import torch

    expected, _, actual, _ = _encode([(string_a, string_b)])
    m, n = len(string_a), len(string_b)
    out = np.zeros((m + 1, n + 1), dtype=np.float32)
    check(lib.aph_edit_matrix(_pointer(expected), m, _pointer(actual), n, _pointer(out)), "aph_edit_matrix")
    return torch.from_numpy(out)


def to_substitutions(string_a: Sequence[str], string_b: Sequence[str], operations: LevenstheinOperations) -> List[Tuple[Action, str, str]]:
    """``src/edit_distance.rs:100-114``."""
    result = []
    for operation, a_index, b_index in operations:
        if operation == Action.DELETION:
            result.append((operation, string_a[a_index], ""))
        elif operation == Action.INSERTION:
            result.append((operation, "", string_b[b_index]))
        else:
            result.append((operation, string_a[a_index], string_b[b_index]))
    return result


class PropertyWeighting:
    """Edit distances whose substitution cost is the number of differing phonetic properties
    (``src/edit_distance.rs:497-598``): ``property_table[symbol]`` is a 1-D tensor / array of properties."""

    def __init__(self, insertion_cost: float, deletion_cost: float, property_table) -> None:
        self._insertion_cost = float(insertion_cost)
        self._deletion_cost = float(deletion_cost)
        self._table = property_table

    def _substitution_costs(self, string_a: Sequence, string_b: Sequence) -> np.ndarray:
        rows_a = np.stack([np.asarray(self._table[symbol]) for symbol in string_a]) if len(string_a) else np.zeros((0, 1))
        rows_b = np.stack([np.asarray(self._table[symbol]) for symbol in string_b]) if len(string_b) else np.zeros((0, 1))
        if not len(string_a) or not len(string_b):
            return np.zeros((max(1, len(string_a) * len(string_b)),), dtype=np.float32)
        return np.ascontiguousarray((rows_a[:, None, :] != rows_b[None, :, :]).sum(-1).astype(np.float32))

    def _run(self, string_a: Sequence, string_b: Sequence, mode: int):
        m, n = len(string_a), len(string_b)
        costs = self._substitution_costs(string_a, string_b)
        matrix = np.zeros((m + 1, n + 1), dtype=np.float32) if mode == 0 else None
        operations = np.zeros((max(1, m + n), 3), dtype=np.int64) if mode == 1 else None
        stats = np.zeros(4, dtype=np.uint64) if mode == 2 else None
        final = ctypes.c_float(0.0)
        pointer = lambda array: None if array is None else _pointer(array)  # noqa: E731
        count = lib.aph_edit_weighted(m, n, _pointer(costs), self._insertion_cost, self._deletion_cost, mode, pointer(matrix), pointer(operations), pointer(stats), ctypes.byref(final))
        if count < 0:
            check(int(count), "aph_edit_weighted")
        return matrix, operations, stats, int(count), float(final.value)

    def levensthein_matrix(self, string_a: Sequence, string_b: Sequence):
        import torch

        return torch.from_numpy(self._run(string_a, string_b, 0)[0])

    def levensthein_operations(self, string_a: Sequence, string_b: Sequence) -> Tuple[LevenstheinOperations, float]:
        _, operations, _, count, cost = self._run(string_a, string_b, 1)
        return [(Action.from_int(int(action)), int(i), int(j)) for action, i, j in operations[:count].tolist()], cost

    def levensthein_statistics(self, string_a: Sequence, string_b: Sequence) -> EditStatistics:
        return EditStatistics(*self._run(string_a, string_b, 2)[2].tolist())


class MissingSegmentError(ValueError):
    """``src/ipa_segmenter.rs:11``."""


def _debug(text: str) -> str:
    """Rust's ``{:?}`` of a ``&str`` for the error message: double quotes, backslash escapes."""
    return '"' + text.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n").replace("\t", "\\t").replace("\r", "\\r") + '"'


class IpaSegmenter:
    """Leftmost-longest segmentation of IPA transcriptions into a known segment vocabulary (``src/ipa_segmenter.rs``)."""

    def __init__(self, ipa_segments: List[str]) -> None:
        self.ipa_segments = list(ipa_segments)
        encoded = [segment.encode("utf-8") for segment in self.ipa_segments]
        offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
        np.cumsum([len(piece) for piece in encoded], out=offsets[1:])
        self._handle = lib.aph_segmenter_create(b"".join(encoded), _pointer(offsets), len(encoded))
        if not self._handle:
            raise RuntimeError("aph_segmenter_create failed")

    def __del__(self) -> None:
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            lib.aph_segmenter_free(handle)

    def _matches(self, word: str) -> Tuple[bytes, List[Tuple[int, int]]]:
        data = word.encode("utf-8")
        bounds = np.zeros((max(1, len(data)), 2), dtype=np.int64)
        count = lib.aph_segmenter_find(self._handle, data, len(data), _pointer(bounds), len(bounds))
        if count < 0:
            check(int(count), "aph_segmenter_find")
        return data, [(int(start), int(end)) for start, end in bounds[:count].tolist()]

    def _segment_word(self, word: str, include_missing: bool) -> List[str]:
        data, matches = self._matches(word)
        pieces: List[str] = []
        last_end = 0
        for start, end in matches:
            if include_missing and start != last_end:
                pieces.append(data[last_end:start].decode("utf-8"))
            pieces.append(data[start:end].decode("utf-8"))
            last_end = end
        if include_missing and last_end != len(data):
            pieces.append(data[last_end:].decode("utf-8"))
        return pieces

    def _segment_word_checked(self, word: str) -> List[str]:
        data, matches = self._matches(word)
        pieces: List[str] = []
        last_end = 0
        for start, end in matches:
            if start != last_end:
                raise MissingSegmentError(f"Segment {_debug(data[last_end:start].decode('utf-8'))} is missing from the vocabulary. Found in: {_debug(word)}")
            pieces.append(data[start:end].decode("utf-8"))
            last_end = end
        if last_end != len(data):
            raise MissingSegmentError(f"Segment {_debug(data[last_end:].decode('utf-8'))} is missing from the vocabulary. Found in: {_debug(word)}")
        return pieces

    def segment(self, transcription: str, include_missing: bool = False) -> List[str]:
        return self._segment_word(transcription, include_missing)

    def segment_checked(self, transcription: str) -> List[str]:
        return self._segment_word_checked(transcription)

    def segment_words(self, transcriped_words: List[str], include_missing: bool = False) -> List[str]:
        return [piece for word in transcriped_words for piece in self._segment_word(word, include_missing)]

    def segment_words_checked(self, transcriped_words: List[str]) -> List[str]:
        return [piece for word in transcriped_words for piece in self._segment_word_checked(word)]
