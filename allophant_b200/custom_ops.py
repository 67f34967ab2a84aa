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

There is no CPU implementation: calling an op with CPU tensors raises.
"""
from __future__ import annotations

from typing import Optional

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


REGISTERED = ("log_softmax", "linear_bf16", "layer_norm", "ctc_nll", "zero_mean_unit_var_norm")
