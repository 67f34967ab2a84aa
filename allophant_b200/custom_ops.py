"""``torch.library`` registration of the C-ABI entry points (namespace ``allophant_b200``).

The engine's launch lists call ``liballophant_b200.so`` through ctypes directly (a dispatcher round trip per launch
would double the host cost of a 200-launch forward pass); the operators a user composes by hand are registered here
as PyTorch custom ops — schema, CUDA implementation, fake (meta) kernel for tracing and, where the reference
differentiates through them, an autograd formula that runs the library's backward kernels:

    torch.ops.allophant_b200.log_softmax(x)                              Allophant.log_probabilities (acoustic_model.py:1051-1052)
    torch.ops.allophant_b200.linear_bf16(x, w, bias, gelu)               nn.Linear call sites (bf16 operands, fp32 accumulate)
    torch.ops.allophant_b200.layer_norm(x, weight, bias, eps)            nn.LayerNorm over the last axis (512 / 1024 columns)
    torch.ops.allophant_b200.ctc_nll(log_probs, labels, in_len, lab_len) per-utterance CTC negative log-likelihood (loss_functions.py:24-27)
    torch.ops.allophant_b200.zero_mean_unit_var_norm(x, lengths)         acoustic_model.py:762-767
    torch.ops.allophant_b200.attention(q, k, v, frame_lengths, heads)    Wav2Vec2Attention's SDPA, key padding from frame counts (HF:466-549, 758-762)
    torch.ops.allophant_b200.ctc_loss(logits, labels, in_len, lab_len)   CTCWrapper.forward: nn.CTCLoss(sum, zero_infinity)(log_softmax(logits)),
                                                                         differentiable w.r.t. the logits (loss_functions.py:19-27)
    torch.ops.allophant_b200.ctc_greedy_decode(log_probs, frame_lengths) GreedyCTCDecoder (predictions.py: collapse repeats, drop blanks)
    torch.ops.allophant_b200.allophone_mapping(phone_logits, matrices, csr_offsets, csr_phones, language_ids)
                                                                         AllophoneMapping.forward (acoustic_model.py:64-120): per-phoneme max

There is no CPU implementation: calling an op with CPU tensors raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops

_NS = "allophant_b200"


@torch.library.custom_op(f"{_NS}::log_softmax", mutates_args=(), device_types="cuda")
def log_softmax(x: Tensor) -> Tensor:
    return ops.log_softmax(x)


@log_softmax.register_fake
def _(x: Tensor) -> Tensor:
    return torch.empty(x.shape, device=x.device, dtype=torch.float32)


def _log_softmax_backward(ctx, grad: Tensor) -> Tensor:
    (out,) = ctx.saved_tensors
    return grad - out.exp() * grad.sum(-1, keepdim=True)


log_softmax.register_autograd(_log_softmax_backward, setup_context=lambda ctx, inputs, output: ctx.save_for_backward(output))


@torch.library.custom_op(f"{_NS}::linear_bf16", mutates_args=(), device_types="cuda")
def linear_bf16(x: Tensor, weight: Tensor, bias: Optional[Tensor], gelu: bool) -> Tensor:
    lead = x.shape[:-1]
    flat = x.reshape(-1, x.shape[-1])
    if flat.dtype != torch.bfloat16:
        flat = ops.cast_bf16(flat.float().contiguous())
    w = weight if weight.dtype == torch.bfloat16 else ops.cast_bf16(weight)
    out = ops.linear_bf16(flat.contiguous(), w.contiguous(), None if bias is None else bias.float().contiguous(), gelu=gelu, out_dtype=torch.float32)
    return out.view(*lead, weight.shape[0])


@linear_bf16.register_fake
def _(x: Tensor, weight: Tensor, bias: Optional[Tensor], gelu: bool) -> Tensor:
    return torch.empty(*x.shape[:-1], weight.shape[0], device=x.device, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::layer_norm", mutates_args=(), device_types="cuda")
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    cols = x.shape[-1]
    flat = x.float().reshape(-1, cols).contiguous()
    out = torch.empty_like(flat)
    ops.layernorm_rows(flat, flat.shape[0], cols, cols, weight.float().contiguous(), bias.float().contiguous(), eps, out_f32=out, ld_f32=cols)
    return out.view(x.shape)


@layer_norm.register_fake
def _(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    return torch.empty(x.shape, device=x.device, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::ctc_nll", mutates_args=(), device_types="cuda")
def ctc_nll(log_probs: Tensor, labels: Tensor, input_lengths: Tensor, label_lengths: Tensor) -> Tensor:
    """``log_probs`` fp32 ``[T', N, classes]`` (time first) -> fp32 ``[N]`` negative log-likelihoods (inf when infeasible)."""
    problem = ops.CtcProblem([log_probs.float()], [labels], [label_lengths], input_lengths, batch_first=False, need_grad=False)
    problem.forward()
    return problem.nll[0].clone()


@ctc_nll.register_fake
def _(log_probs: Tensor, labels: Tensor, input_lengths: Tensor, label_lengths: Tensor) -> Tensor:
    return torch.empty(log_probs.shape[1], device=log_probs.device, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::zero_mean_unit_var_norm", mutates_args=(), device_types="cuda")
def zero_mean_unit_var_norm(features: Tensor, lengths: Tensor) -> Tensor:
    return ops.zero_mean_unit_var_norm(features, lengths)


@zero_mean_unit_var_norm.register_fake
def _(features: Tensor, lengths: Tensor) -> Tensor:
    return torch.empty(features.shape, device=features.device, dtype=torch.float32)


@torch.library.custom_op(f"{_NS}::attention", mutates_args=(), device_types="cuda")
def attention(q: Tensor, k: Tensor, v: Tensor, frame_lengths: Tensor, heads: int) -> Tensor:
    """``softmax(q k^T / sqrt(64)) v`` per (utterance, head): ``q`` / ``k`` / ``v`` bf16 ``[n_utt * heads, frames, 64]``,
    ``frame_lengths`` ``[n_utt]`` -> context bf16 ``[n_utt * frames, heads * 64]``; keys at or past an utterance's frame count get
    probability 0.  (The kernel takes queries pre-scaled by ``log2(e) / 8`` — the encoder's QKV epilogue does that in fp32 —
    and exponentiates in base 2; this op applies the scale itself.)"""
    if q.dtype != torch.bfloat16 or q.dim() != 3 or q.shape[-1] != 64 or q.shape[0] % heads != 0:
        raise ValueError("attention expects bf16 [n_utt * heads, frames, 64] operands")
    n_utt, seq = q.shape[0] // heads, q.shape[1]
    ctx = torch.empty(n_utt * seq, heads * 64, device=q.device, dtype=torch.bfloat16)
    scaled = (q.float() * (1.4426950408889634 / 8.0)).bfloat16().contiguous()
    ops.attention(scaled, k.contiguous(), v.contiguous(), ctx, frame_lengths.to(torch.int32).contiguous(), n_utt, heads, seq)
    return ctx


@attention.register_fake
def _(q: Tensor, k: Tensor, v: Tensor, frame_lengths: Tensor, heads: int) -> Tensor:
    return torch.empty((q.shape[0] // heads) * q.shape[1], heads * 64, device=q.device, dtype=torch.bfloat16)


@torch.library.custom_op(f"{_NS}::ctc_loss_with_gradient", mutates_args=(), device_types="cuda")
def ctc_loss_with_gradient(logits: Tensor, labels: Tensor, input_lengths: Tensor, label_lengths: Tensor) -> Tuple[Tensor, Tensor]:
    """(sum over the batch of the CTC negative log-likelihoods of ``log_softmax(logits)`` with infinite losses zeroed — fp32
    ``[1]`` —, its gradient w.r.t. ``logits``): the alpha and the fused beta / gradient launch of the training step."""
    log_probs = ops.log_softmax(logits.float())
    problem = ops.CtcProblem([log_probs], [labels], [label_lengths], input_lengths, batch_first=False, need_grad=True)
    loss = problem.forward().clone()
    (gradient,) = problem.backward(torch.ones(1, device=logits.device))
    return loss, gradient.contiguous()


@ctc_loss_with_gradient.register_fake
def _(logits: Tensor, labels: Tensor, input_lengths: Tensor, label_lengths: Tensor) -> Tuple[Tensor, Tensor]:
    return torch.empty(1, device=logits.device, dtype=torch.float32), torch.empty(logits.shape, device=logits.device, dtype=torch.float32)


def _ctc_backward(ctx, grad_loss: Tensor, _grad_gradient: Optional[Tensor]):
    (gradient,) = ctx.saved_tensors
    return gradient * grad_loss.reshape(1, 1, 1), None, None, None


ctc_loss_with_gradient.register_autograd(_ctc_backward, setup_context=lambda ctx, inputs, output: ctx.save_for_backward(output[1]))


def ctc_loss(logits: Tensor, labels: Tensor, input_lengths: Tensor, label_lengths: Tensor) -> Tensor:
    """``CTCWrapper.forward`` on one head through the custom op; ``logits`` fp32 ``[T', N, classes]`` (time first)."""
    return torch.ops.allophant_b200.ctc_loss_with_gradient(logits, labels, input_lengths, label_lengths)[0][0]


@torch.library.custom_op(f"{_NS}::ctc_greedy_decode", mutates_args=(), device_types="cuda")
def ctc_greedy_decode(log_probs: Tensor, frame_lengths: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """``log_probs`` fp32 ``[T', N, classes]`` (time first, blank = 0) -> (tokens int32 ``[N, T']`` — the first ``counts[n]`` of a
    row are the hypothesis —, counts int32 ``[N]``, scores fp32 ``[N]`` = sum of the frame-wise maxima)."""
    seq, n_utt, classes = log_probs.shape
    flat = log_probs.float().transpose(0, 1).contiguous().view(n_utt * seq, classes)
    best = torch.empty(n_utt * seq, device=flat.device, dtype=torch.int32)
    best_value = torch.empty(n_utt * seq, device=flat.device, dtype=torch.float32)
    ops.argmax_rows(flat, classes, n_utt * seq, classes, best, best_value)
    tokens, _, counts, scores = ops.ctc_greedy_collapse(best, best_value, frame_lengths.to(torch.int32).contiguous(), n_utt, seq, n_utt, 0)
    return tokens, counts, scores


@ctc_greedy_decode.register_fake
def _(log_probs: Tensor, frame_lengths: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    seq, n_utt, _ = log_probs.shape
    device = log_probs.device
    return (torch.empty(n_utt, seq, device=device, dtype=torch.int32), torch.empty(n_utt, device=device, dtype=torch.int32),
            torch.empty(n_utt, device=device, dtype=torch.float32))  # fmt: skip


@torch.library.custom_op(f"{_NS}::allophone_mapping", mutates_args=(), device_types="cuda")
def allophone_mapping(phone_logits: Tensor, matrices: Tensor, csr_offsets: Tensor, csr_phones: Tensor, language_ids: Tensor) -> Tensor:
    """``phone_logits`` fp32 ``[N, T', P + 1]`` -> phoneme logits ``[N, T', Q + 1]``: for phoneme q of utterance n's language the maximum
    over its allophones p of ``logits[p] * matrices[language, p, q]``; (``csr_offsets``, ``csr_phones``) list the allophones of every
    (language, phoneme) pair (``HeadsRuntime._allophone_csr``)."""
    mapped, _ = ops.allophone_forward(phone_logits.float(), matrices.float().contiguous(), csr_offsets, csr_phones, language_ids.to(torch.int64).contiguous())
    return mapped


@allophone_mapping.register_fake
def _(phone_logits: Tensor, matrices: Tensor, csr_offsets: Tensor, csr_phones: Tensor, language_ids: Tensor) -> Tensor:
    return torch.empty(phone_logits.shape[0], phone_logits.shape[1], matrices.shape[2], device=phone_logits.device, dtype=torch.float32)


REGISTERED = ("log_softmax", "linear_bf16", "layer_norm", "ctc_nll", "zero_mean_unit_var_norm", "attention", "ctc_loss_with_gradient",
              "ctc_greedy_decode", "allophone_mapping")  # fmt: skip
