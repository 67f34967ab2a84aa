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
    INSERTION = 1
    DELETION = 2
    SUBSTITUTION = 3

    @staticmethod
    def from_int(integer: int) -> "Action":
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
