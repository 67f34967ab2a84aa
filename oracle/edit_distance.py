"""TEST INFRASTRUCTURE — ctypes access to the C restatement of the reference's Rust edit distance
(``oracle/edit_distance.c``; follows ``src/edit_distance.rs``)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_edit.so")


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    lib = ctypes.CDLL(_LIB_PATH)
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib.ora_levenshtein.argtypes = [i64p, ctypes.c_int64, i64p, ctypes.c_int64]
    lib.ora_levenshtein.restype = ctypes.c_uint64
    lib.ora_levenshtein_statistics.argtypes = [i64p, ctypes.c_int64, i64p, ctypes.c_int64, ctypes.POINTER(ctypes.c_uint64)]
    lib.ora_levenshtein_statistics.restype = None
    lib.ora_levenshtein_operations.argtypes = [i64p, ctypes.c_int64, i64p, ctypes.c_int64, i64p, i64p]
    lib.ora_levenshtein_operations.restype = ctypes.c_float
    lib.ora_word_error_rate.argtypes = [ctypes.c_uint64] * 4
    lib.ora_word_error_rate.restype = ctypes.c_float
    return lib


_lib = _load()


def _array(values: Sequence[int]):
    return (ctypes.c_int64 * max(1, len(values)))(*values)


def levenshtein(a: Sequence[int], b: Sequence[int]) -> int:
    return int(_lib.ora_levenshtein(_array(a), len(a), _array(b), len(b)))


def statistics(a: Sequence[int], b: Sequence[int]) -> Tuple[int, int, int, int]:
    """(insertions, deletions, substitutions, correct)."""
    out = (ctypes.c_uint64 * 4)()
    _lib.ora_levenshtein_statistics(_array(a), len(a), _array(b), len(b), out)
    return tuple(int(v) for v in out)


def operations(a: Sequence[int], b: Sequence[int]) -> Tuple[List[Tuple[int, int, int]], float]:
    ops = (ctypes.c_int64 * (3 * (len(a) + len(b) + 1)))()
    count = ctypes.c_int64(0)
    cost = _lib.ora_levenshtein_operations(_array(a), len(a), _array(b), len(b), ops, ctypes.byref(count))
    return [(int(ops[3 * i]), int(ops[3 * i + 1]), int(ops[3 * i + 2])) for i in range(count.value)], float(cost)


def word_error_rate(insertions: int, deletions: int, substitutions: int, correct: int) -> float:
    return float(_lib.ora_word_error_rate(insertions, deletions, substitutions, correct))
