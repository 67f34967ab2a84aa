"""Padding arithmetic of the convolutional front-end layers (mirrors ``allophant/network/padding.py:10-21``)."""
from __future__ import annotations

from typing import Tuple


def get_padding(kernel_size: int, stride: int = 1, stft_type: bool = False) -> Tuple[int, int]:
    """(left, right) padding of a 1-D filter: "same"-style for stride 1, enough right padding for a strided filter to
    reach the edge otherwise."""
    half = kernel_size // 2
    if stft_type:
        return (half, half - 1) if stride == 1 else (half, half)
    if stride > 1:
        return (half, kernel_size - 1)
    return (half, half)
