"""Data-parallel plumbing: one process per GPU, utterances sharded across ranks.

The reference has no distributed code.  The hot path shards by utterance (no
cross-utterance op exists), so inference needs NO collective on the data path: every rank runs
``Estimator.predict`` on its own shard and only the decoded hypotheses (host objects) are gathered.
Training adds the two real exchange steps of §8e: a sum all-reduce of the gradient buckets and of
the scalar loss normaliser (``estimator.py:737`` divides by the label count of the WHOLE batch).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .dataset_processing import Batch


def shard_indices(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced assignment: utterances sorted by length (longest first), dealt round-robin in
    serpentine order so every rank gets a similar number of samples and padding stays small."""
    order = sorted(range(len(lengths)), key=lambda index: (-int(lengths[index]), index))
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for position, index in enumerate(order):
        round_index, slot = divmod(position, world_size)
        rank = slot if round_index % 2 == 0 else world_size - 1 - slot
        shards[rank].append(index)
    return shards


def shard_batch(batch: Batch, rank: int, world_size: int) -> Tuple[Batch, List[int]]:
    """The sub-batch of ``rank`` (re-padded to its own longest utterance) and the original indices."""
    lengths = batch.lengths.tolist()
    indices = shard_indices(lengths, world_size)[rank]
    if not indices:
        return Batch(batch.audio_features[:0], batch.lengths[:0], batch.language_ids[:0]), indices
    select = torch.tensor(indices, device=batch.lengths.device)
    longest = max(lengths[i] for i in indices)
    return (
        Batch(
            batch.audio_features.index_select(0, select)[:, :longest].contiguous(),
            batch.lengths.index_select(0, select),
            batch.language_ids.index_select(0, select),
        ),
        indices,
    )


def gather_by_index(local: Dict[int, Any], group: Optional[dist.ProcessGroup] = None) -> Dict[int, Any]:
    """Gathers per-utterance host results (e.g. decoded hypotheses) from every rank, keyed by original index."""
    if not dist.is_available() or not dist.is_initialized():
        return dict(local)
    gathered: List[Optional[Dict[int, Any]]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, local, group=group)
    merged: Dict[int, Any] = {}
    for part in gathered:
        merged.update(part or {})
    return merged


def allreduce_gradients(parameters: Iterable[Tensor], bucket_bytes: int = 48 << 20, group: Optional[dist.ProcessGroup] = None) -> int:
    """Sum all-reduce of ``.grad`` in flat buckets (≈48 MB: a few launches, each far above NCCL's latency floor).
    Returns the number of collectives issued."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    grads = [p.grad for p in parameters if p.grad is not None]
    issued, bucket, size = 0, [], 0

    def flush() -> None:
        nonlocal issued, bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        offset = 0
        for g in bucket:
            g.copy_(flat[offset : offset + g.numel()].view_as(g))
            offset += g.numel()
        issued += 1
        bucket, size = [], 0

    for grad in grads:
        bucket.append(grad)
        size += grad.numel() * grad.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return issued


class GradientReducer:
    """Sum all-reduce of gradient groups, overlapped with the backward pass that produces them.

    The CUDA backward (``EncoderPlan.backward``) writes the gradients of one encoder layer into one flat
    buffer and calls ``submit`` as soon as that layer's kernels are enqueued: the collective of layer ``i``
    (≈50 MB for XLS-R-300M, far above NCCL's latency floor, one ring/NVLS pass over NVLink) runs on NCCL's
    stream while layers ``i-1 … 0`` are still computing.  ``finish`` makes the compute stream wait for all of
    them before autograd accumulates the (views of the) buffers into ``.grad``.  The gradients are SUMMED:
    divide the loss by the global label count (``global_label_count``) to reproduce the reference's
    single-device arithmetic (``estimator.py:737``) exactly."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, wire_dtype: Optional[torch.dtype] = None) -> None:
        """``wire_dtype=torch.bfloat16`` sends the gradients as bf16 (half the bytes over NVLink; the sum then carries bf16
        rounding, so the default keeps fp32 and the reference's single-device arithmetic)."""
        self.group = group
        self.wire_dtype = wire_dtype
        self.works: List[Any] = []
        self._compressed: List[Tuple[Tensor, Tensor]] = []
        self.issued = 0
        self.bytes = 0

    @property
    def active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def submit(self, flat: Tensor, views: Optional[Dict[str, Tensor]] = None) -> None:
        if not self.active:
            return
        wire = flat
        if self.wire_dtype is not None and flat.dtype != self.wire_dtype:
            wire = flat.to(self.wire_dtype)
            self._compressed.append((flat, wire))
        self.works.append(dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.issued += 1
        self.bytes += wire.numel() * wire.element_size()

    def submit_tensors(self, named: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """Packs small tensors (the classifier heads' gradients) into one bucket; returns views of the bucket."""
        if not self.active or not named:
            return named
        flat = torch.cat([value.reshape(-1) for value in named.values()])
        views, offset = {}, 0
        for name, value in named.items():
            views[name] = flat[offset : offset + value.numel()].view(value.shape)
            offset += value.numel()
        self.submit(flat)
        return views

    def finish(self) -> None:
        for work in self.works:
            work.wait()
        self.works.clear()
        for flat, wire in self._compressed:  # back into the fp32 buffers the .grad tensors are views of
            flat.copy_(wire)
        self._compressed.clear()


def attach_gradient_reducer(model: Any, reducer: Optional[GradientReducer]) -> None:
    """Makes ``model`` (an ``Allophant``) all-reduce its gradients inside ``backward()`` (``None`` detaches)."""
    model._heads.gradient_reducer = reducer


def global_label_count(local_label_lengths: Sequence[Tensor], group: Optional[dist.ProcessGroup] = None) -> Tensor:
    """Σ label lengths over all heads and all ranks — the divisor of the step loss (``estimator.py:737``)."""
    total = torch.cat([lengths.reshape(-1) for lengths in local_label_lengths]).sum().to(torch.float64)  # two launches, not one per head
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total
