"""Host helpers on the hot path (mirrors ``allophant/utils.py:45-76``)."""
from __future__ import annotations

import typing
from contextlib import contextmanager

import torch
from torch import Tensor, nn


def mask_sequence(
    lengths: Tensor, max_length: "int | None" = None, start: int = 0, inverse: bool = False, batch_first: bool = True
) -> Tensor:
    """Boolean ``batch x max length`` mask of the valid positions of variable-length sequences.

    Same semantics as the reference: ``max_length`` defaults to ``lengths.max()`` (a host
    sync — the CUDA kernels of this package take ``lengths`` directly and never build this
    mask; the function exists for API parity and host-side code).
    """
    if max_length is None:
        max_length = typing.cast(int, int(lengths.max()))
    positions = torch.arange(start, max_length, device=lengths.device)
    if batch_first:
        positions, bounds = positions.unsqueeze(0), lengths.unsqueeze(1)
    else:
        positions, bounds = positions.unsqueeze(1), lengths.unsqueeze(0)
    return positions >= bounds if inverse else positions < bounds


@contextmanager
def evaluation(model: nn.Module):
    """Temporarily switches a module to eval mode (``estimator.py:138-150``)."""
    was_training = model.training
    if not was_training:
        # already in eval mode (the usual case for inference): walking ~4000 sub-modules twice per call to set
        # flags that are already set would cost more host time than enqueueing the whole forward pass
        yield model
        return
    model.eval()
    try:
        yield model
    finally:
        model.train(was_training)
