"""Length arithmetic of the convolutional front end (mirrors ``allophant/network/frontend.py:192-203``)."""
from __future__ import annotations

from typing import Callable

import torch
from torch import Tensor


def conv_length(kernel_size: int, stride: int = 1, use_padding: bool = True, stft_type: bool = False) -> Callable[[Tensor], Tensor]:
    """Output length of a 1-D convolution.  Only the unpadded form used by the wav2vec2 feature
    extractor (``acoustic_model.py:823-826``) is provided."""
    if use_padding:
        raise NotImplementedError("padded frontends belong to the from-scratch transformer encoder (not in this build)")

    def padded_length(lengths: Tensor) -> Tensor:
        return torch.div(lengths - kernel_size, stride, rounding_mode="floor") + 1

    return padded_length
