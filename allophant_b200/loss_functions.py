"""CTC loss (drop-in for ``allophant/loss_functions.py:19-27``) backed by the multi-head CUDA kernels."""
from __future__ import annotations

from abc import ABCMeta, abstractmethod
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor, nn

from . import ops


class LossWrapper(nn.Module, metaclass=ABCMeta):
    @abstractmethod
    def forward(self, logits: Tensor, labels: Tensor, predicted_lengths: Tensor, label_lengths: Tensor) -> Tensor:
        pass


class _MultiHeadCtc(torch.autograd.Function):
    """sum-reduced CTC negative log-likelihood of H heads; gradient w.r.t. the LOGITS."""

    @staticmethod
    def forward(ctx, labels, label_lengths, input_lengths, *logits):
        need_grad = any(t.requires_grad for t in logits)
        log_probs = ops.log_softmax_many([t.detach() for t in logits])  # one launch for all heads
        problem = ops.CtcProblem(log_probs, labels, label_lengths, input_lengths, batch_first=False, need_grad=need_grad)
        loss = problem.forward()
        ctx.problem = problem
        ctx.n = len(logits)
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_loss):
        grads = ctx.problem.backward(grad_loss.contiguous())
        return (None, None, None, *grads)


def multi_head_ctc_loss(
    logits: Sequence[Tensor], labels: Sequence[Tensor], input_lengths: Tensor, label_lengths: Sequence[Tensor]
) -> Tensor:
    """fp32 ``[H]``: for each head ``nn.CTCLoss(reduction="sum", zero_infinity=True)(log_softmax(logits_h), ...)``.

    ``logits[h]`` is time-first ``[T', N, classes_h]`` (any strides with a contiguous class axis),
    ``labels[h]`` int64 ``[N, S_max_h]``, ``label_lengths[h]`` int64 ``[N]``.  One alpha launch and one
    beta launch cover all heads."""
    return _MultiHeadCtc.apply(list(labels), list(label_lengths), input_lengths, *logits)


class CTCWrapper(LossWrapper):
    def __init__(self):
        super().__init__()

    def forward(self, logits: Tensor, labels: Tensor, predicted_lengths: Tensor, label_lengths: Tensor) -> Tensor:
        return multi_head_ctc_loss([logits], [labels], predicted_lengths, [label_lengths])[0]
