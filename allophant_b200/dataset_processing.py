"""Batch containers: the input contract of the hot path.

Field-for-field mirrors of ``allophant/dataset_processing.py`` ``Batch`` (49-85) and
``LabeledBatch`` (132-162).  ``audio_features`` is fp32 ``[N, T]`` zero-padded to exactly
``lengths.max()`` samples, ``lengths`` / ``language_ids`` are int64 ``[N]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Dict, Iterator, List, Tuple

import torch
from torch import Tensor


@dataclass
class Batch:
    audio_features: Tensor
    lengths: Tensor
    language_ids: Tensor

    def pin_memory(self):
        self.audio_features = self.audio_features.pin_memory()
        self.lengths = self.lengths.pin_memory()
        self.language_ids = self.language_ids.pin_memory()
        return self

    def _inputs_to(self, device, non_blocking: bool = False, copy: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
        return (
            self.audio_features.to(device, non_blocking=non_blocking, copy=copy),
            self.lengths.to(device, non_blocking=non_blocking, copy=copy),
            self.language_ids.to(device, non_blocking=non_blocking, copy=copy),
        )

    def to(self, device, non_blocking: bool = False, copy: bool = False):
        return self.__class__(*self._inputs_to(device, non_blocking, copy))

    def cuda(self, non_blocking: bool = False, copy: bool = False):
        return self.to("cuda", non_blocking, copy)

    def size(self) -> int:
        return len(self)

    def __len__(self) -> int:
        return self.lengths.numel()

    def __repr__(self) -> str:
        return "{}(Features: ({}; {}))".format(
            self.__class__.__name__, self.audio_features.shape, self.audio_features.dtype
        )


@dataclass(repr=False)
class LabeledBatch(Batch):
    """``attribute_indices``: one dict per G2P engine, name -> int64 ``[N, S_max(name)]`` (zero padded,
    labels >= 1); ``label_lengths``: one int64 ``[n_features, N]`` per engine; ``label_length_indices``:
    name -> row of ``label_lengths``."""

    attribute_indices: List[Dict[str, Tensor]]
    label_lengths: List[Tensor]
    label_length_indices: Dict[str, int]

    def pin_memory(self):
        super().pin_memory()
        self.attribute_indices = [{k: v.pin_memory() for k, v in d.items()} for d in self.attribute_indices]
        self.label_lengths = [l.pin_memory() for l in self.label_lengths]
        return self

    def to(self, device, non_blocking: bool = False, copy: bool = False):
        return self.__class__(
            *self._inputs_to(device, non_blocking, copy),
            [{k: v.to(device, non_blocking=non_blocking, copy=copy) for k, v in d.items()} for d in self.attribute_indices],
            [l.to(device, non_blocking=non_blocking, copy=copy) for l in self.label_lengths],
            label_length_indices=self.label_length_indices,
        )


@dataclass(repr=False)
class RawLabeledBatch(Batch):
    """``dataset_processing.py:92-130``: audio with raw (string) transcriptions per G2P engine and utterance ids."""

    raw_labels: List[List[List[str]]]
    utterance_ids: List[str]

    def to(self, device, non_blocking: bool = False, copy: bool = False):
        return self.__class__(*self._inputs_to(device, non_blocking, copy), self.raw_labels, self.utterance_ids)

    def split_by_language(self) -> Iterator[Tuple[int, "RawLabeledBatch"]]:
        """One sub-batch per run of consecutive equal language ids (``dataset_processing.py:103-130``), each re-padded to its
        own longest utterance; yields ``(language id, batch)``."""
        languages, run_lengths = self.language_ids.unique_consecutive(return_counts=True)
        start = 0
        for language, run in zip(languages, run_lengths.tolist()):
            stop = start + run
            lengths = self.lengths[start:stop]
            yield (
                language,
                type(self)(
                    self.audio_features[start:stop, ..., : int(lengths.max())],
                    lengths,
                    self.language_ids[start:stop],
                    [engine_labels[start:stop] for engine_labels in self.raw_labels],
                    self.utterance_ids[start:stop],
                ),
            )
            start = stop


class BatchType(Enum):
    """``dataset_processing.py:165-173``."""

    UNLABELED = 0
    RAW = 1
    INDEXED = 2
