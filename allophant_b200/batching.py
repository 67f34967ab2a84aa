"""Host feeding (drop-in for the batch assembly of ``allophant/batching.py``).

* ``MaxFrameBatchSampler`` / ``SkipBatchSampler`` (``batching.py:94-168``): frame-budget batching, same iteration
  behaviour (including the empty first batch the reference yields when the very first utterance exceeds the budget).
* ``build_batch`` (``batching.py:171-215``, ``_build_batch``): merges single-entry batches into one dense ``Batch`` /
  ``LabeledBatch`` / ``RawLabeledBatch``.  The audio is collated by ``aph_collate_pad_f32`` (host C++ threads) straight
  into a PINNED buffer the batch owns, so the host -> device copy that follows is a single asynchronous
  DMA from page-locked memory (``torch.nn.utils.rnn.pad_sequence`` + pageable ``.to(device)`` costs two extra passes over
  the batch and a synchronous copy).
* ``shard_for_rank``: per-rank, length-balanced sub-batches for data-parallel runs (``distributed.shard_indices``).
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence

import torch
from torch import Tensor
from torch.nn.utils import rnn
from torch.utils.data import BatchSampler, Sampler

from ._lib import check, lib
from .dataset_processing import Batch, BatchType, LabeledBatch, RawLabeledBatch  # noqa: F401


class MaxFrameBatchSampler(BatchSampler):
    """Batches indices until (batch size x longest sequence) would exceed ``batch_size`` frames (``batching.py:94-139``)."""

    def __init__(self, sampler: "Sampler[int] | Iterable[int]", batch_size: int, frame_lengths: Tensor) -> None:
        self._sampler = sampler
        self._batch_size = batch_size
        self._frame_lengths = frame_lengths

    def __iter__(self) -> Iterator[List[int]]:
        batch_indices: List[int] = []
        max_length = 0
        for index in self._sampler:
            length = self._frame_lengths[index]
            if length > max_length:
                max_length = length
            new_batch_size = (len(batch_indices) + 1) * max_length
            if new_batch_size > self._batch_size:
                yield batch_indices
                max_length = length
                batch_indices = [index]
            else:
                batch_indices.append(index)
        if batch_indices:
            yield batch_indices


class SkipBatchSampler(BatchSampler):
    """Skips the first ``skip_count`` batches of another batch sampler (``batching.py:142-161``)."""

    def __init__(self, sampler: BatchSampler, skip_count: int) -> None:
        self._sampler = sampler
        self._skip_count = skip_count

    def __iter__(self) -> Iterator[List[int]]:
        samples = iter(self._sampler)
        for _, _ in zip(samples, range(self._skip_count)):
            pass
        return samples


def _pinned_staging(count: int, longest: int) -> Tensor:
    """An OWNED page-locked fp32 ``[count, longest]`` tensor for one batch.  It comes from torch's caching host allocator, which
    recycles a block only after the tensor died AND every asynchronous copy recorded on it finished — so any number of
    batches can be alive at once (the reference's ``_training_batch_accumulation`` holds ``accumulation_factor`` host
    batches before ``.to(device)``; a prefetching consumer holds more) and none aliases another.  Inside a forked
    ``DataLoader`` worker (CUDA must not be initialised there) and without a GPU the buffer is pageable."""
    pin = torch.cuda.is_available() and torch.utils.data.get_worker_info() is None
    with torch.inference_mode(False):
        return torch.empty(count, longest, dtype=torch.float32, pin_memory=pin)


def collate_audio(utterances: Sequence[Tensor], pinned: bool = True, n_threads: int = 0) -> Tensor:
    """``rnn.pad_sequence(utterances, batch_first=True)`` for 1-D fp32 host tensors, written by host threads into a
    pinned staging buffer (``pinned=False``: a fresh pageable tensor)."""
    count = len(utterances)
    lengths = [int(u.shape[0]) for u in utterances]
    longest = max(lengths) if lengths else 0
    sources = [u if (u.dtype == torch.float32 and u.is_contiguous()) else u.float().contiguous() for u in utterances]
    if any(u.is_cuda for u in sources):
        raise RuntimeError("collate_audio assembles HOST batches; move the finished batch to the GPU with Batch.to")
    if pinned:
        out = _pinned_staging(count, longest)
    else:
        out = torch.empty(count, longest, dtype=torch.float32)
    if count == 0 or longest == 0:
        return out
    pointers = (ctypes.c_void_p * count)(*[u.data_ptr() for u in sources])
    sizes = (ctypes.c_int64 * count)(*lengths)
    check(lib.aph_collate_pad_f32(pointers, sizes, count, longest, out.data_ptr(), n_threads), "aph_collate_pad_f32")
    return out


def build_batch(batch_type: BatchType, pinned: bool = True) -> Callable[[Sequence[Batch]], Batch]:
    """``_build_batch`` (``batching.py:171-215``): a collate function for single-entry batches."""

    def _create_batch(entries: Sequence[Batch]) -> Batch:
        lengths = torch.tensor([int(entry.lengths) for entry in entries], dtype=torch.long)
        language_ids = torch.tensor([int(entry.language_ids) for entry in entries], dtype=torch.long)
        first = entries[0].audio_features if entries else None
        if first is not None and first.dim() == 1:
            audio_features = collate_audio([entry.audio_features for entry in entries], pinned)
        else:  # feature matrices [frames, features]: the reference's generic path (not used by the wav2vec2 front end)
            audio_features = rnn.pad_sequence([entry.audio_features for entry in entries], True)
            if audio_features.ndim > 2:
                audio_features = audio_features.transpose(1, 2)
        if batch_type == BatchType.UNLABELED:
            return Batch(audio_features, lengths, language_ids)
        if batch_type == BatchType.RAW:
            return RawLabeledBatch(
                audio_features,
                lengths,
                language_ids,
                [list(labels) for labels in zip(*(entry.raw_labels[0] for entry in entries))],
                [entry.utterance_ids[0] for entry in entries],
            )
        label_lengths: List[Tensor] = []
        attribute_indices: List[Dict[str, Tensor]] = []
        if entries:
            for engine in range(len(entries[0].attribute_indices)):
                label_lengths.append(torch.stack([entry.label_lengths[engine] for entry in entries], 1))
                attribute_indices.append(
                    {
                        indices[0][0]: rnn.pad_sequence([tensor for _, tensor in indices], True)
                        for indices in zip(*(entry.attribute_indices[engine].items() for entry in entries))
                    }
                )
        return LabeledBatch(
            audio_features, lengths, language_ids, attribute_indices, label_lengths, entries[0].label_length_indices if entries else {}
        )

    return _create_batch


def shard_for_rank(batch: Batch, rank: int, world_size: int) -> Batch:
    """The length-balanced sub-batch of ``rank`` (see ``distributed.shard_batch``), re-padded to its own longest utterance."""
    from .distributed import shard_batch

    return shard_batch(batch, rank, world_size)[0]
