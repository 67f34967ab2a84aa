"""TEST INFRASTRUCTURE — CPU restatement (plain torch, fp32, differentiable) of the reference's from-scratch transformer
acoustic model: ``TransformerAcousticModel.forward`` (``allophant/network/acoustic_model.py:669-691``) with
``LinearFrontend`` / ``DirectFrontend`` (``frontend.py:154-189``), ``SequentialFrontend`` of ``Glu1d`` / LayerNorm / Dropout layers
(``frontend.py:49-136, 219-276``, ``padding.py:24-53`` including its batch-0 left reflection), ``SinusoidalPositionEmbeddings``
(``acoustic_model.py:34-69``) and ``PreLMTransformerEncoderLayer`` (``acoustic_model.py:281-329``; ``nn.MultiheadAttention``
arithmetic written out).  Weights come from a ``state_dict`` with the reference's keys.

Pinned: ``tests/test_transformer_model.py`` checks it against the hidden states the UNMODIFIED reference produced
(``tests/golden/transformer_*.pt``, made by ``oracle/make_golden_transformer.py``).  It travels to the GPU box, where its
autograd gradients are the reference for the CUDA backward pass.  Nothing in ``allophant_b200`` imports this module."""
from __future__ import annotations

import math
from typing import Any, Dict, List, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def _glu1d(h: Tensor, lengths: Tensor, weight: Tensor, bias: Tensor, kernel: int, stride: int) -> Tuple[Tensor, Tensor]:
    """LengthWrapper mask -> VariableLengthReflectPad -> Conv1d -> GLU on channels-last ``h`` [N, L, C]."""
    n_utt, length, _ = h.shape
    left, right = (kernel // 2, kernel - 1) if stride > 1 else (kernel // 2, kernel // 2)
    masked = h * (torch.arange(length)[None, :] < lengths[:, None])[..., None]
    padded = F.pad(masked, (0, 0, left, right))
    # padding.py:44-46: the index tensor has batch size 1, so EVERY utterance gets utterance 0's left reflection
    if left > 0:
        padded[:, :left] = masked[0, torch.arange(left, 0, -1)][None]
    rows = []
    for n in range(n_utt):
        row, size = padded[n], int(lengths[n])
        if right > 0:
            source = size - 2 - torch.arange(right)
            row = row.index_put((size + left + torch.arange(right),), masked[n, source])
        rows.append(row)
    padded = torch.stack(rows)
    out = F.conv1d(padded.transpose(1, 2), weight, bias, stride=stride).transpose(1, 2)
    half = weight.shape[0] // 2
    return out[..., :half] * torch.sigmoid(out[..., half:]), torch.div(lengths + left + right - kernel, stride, rounding_mode="floor") + 1


def forward(
    state: Mapping[str, Tensor], options: Mapping[str, Any], features: Tensor, lengths: Tensor, prefix: str = "_acoustic_model.",
    masks: Optional[Mapping[str, Tensor]] = None,
) -> Tuple[List[Tensor], Tensor]:  # fmt: skip
    """``options``: the ``acoustic`` dict of a golden case (``transformer`` / ``frontend`` / ``sequential_frontend`` /
    ``elementwise_affine``).  Returns every layer's output after the final LayerNorm, batch-first ``[N, L', d]``, and the frame
    counts.  ``masks``: the reference's train()-mode dropout layers as EXPLICIT multiplicative masks (keep / (1-p) or 0),
    batch-first: ``frontend_input`` [N, L, F] (frontend.py:161-164, 173-180), ``input`` (acoustic_model.py:673),
    ``sequential.<position>`` (frontend.py:236-237), ``attention.<l>`` [N, heads, T, T], ``attention_output.<l>`` (dropout1),
    ``activation.<l>`` (dropout), ``feed_forward_output.<l>`` (dropout2) (acoustic_model.py:323-328)."""
    get = lambda name: state.get(prefix + name)  # noqa: E731
    masks = masks or {}
    drop = lambda value, key: value * masks[key] if key in masks else value  # noqa: E731
    h = drop(features.transpose(1, 2), "frontend_input")  # [N, L, F]
    frontend = options["frontend"]
    if frontend["architecture"] == "linear":
        shift = 1 if frontend.get("input_dropout", 0) > 0 else 0  # nn.Sequential: [Dropout,] LayerNorm, Linear, LeakyReLU
        h = F.layer_norm(h, (h.shape[-1],), get(f"_frontend._layer.{shift}.weight"), get(f"_frontend._layer.{shift}.bias"))
        h = F.leaky_relu(h @ get(f"_frontend._layer.{shift + 1}.weight").T + get(f"_frontend._layer.{shift + 1}.bias"))
    h = drop(h, "input")
    for index, layer in enumerate(options.get("sequential_frontend") or []):
        base = f"_sequential_frontend._layers.layers.{index}.module."
        if layer["type"] == "glu1d":
            h, lengths = _glu1d(h, lengths, get(base + "_weights.weight"), get(base + "_weights.bias"), layer["kernel"], layer.get("stride", 1))
        elif layer["type"] == "layer_norm":
            h = F.layer_norm(h, (h.shape[-1],), get(base + "1.weight"), get(base + "1.bias"))
        elif layer["type"] == "dropout":
            h = drop(h, f"sequential.{index}")
        else:
            raise NotImplementedError(layer["type"])
    n_utt, frames, width = h.shape
    transformer = options["transformer"]
    if transformer.get("positional_embeddings", True):
        component = torch.exp(torch.arange(0, width, 2, dtype=torch.float) * -(math.log(10000) / width))
        bases = torch.stack([component] * 2, 1).view(-1)
        positions = torch.arange(frames, dtype=torch.float)[:, None] * bases
        positions = torch.stack([torch.sin(positions[:, 0::2]), torch.cos(positions[:, 1::2])], -1).view(frames, width)
        h = h + positions[None]
    heads = transformer["heads"]
    activation = F.gelu if transformer.get("activation", "relu") == "gelu" else F.relu
    padding = torch.arange(frames)[None, :] >= lengths[:, None]
    outputs = []
    for index in range(transformer.get("num_layers", 1)):
        base = f"_transformer.layers.{index}."
        src = F.layer_norm(h, (width,), get(base + "norm1.weight"), get(base + "norm1.bias"))
        q, k, v = (src @ get(base + "self_attn.in_proj_weight").T + get(base + "self_attn.in_proj_bias")).split(width, -1)
        split = lambda t: t.view(n_utt, frames, heads, width // heads).transpose(1, 2)  # noqa: E731
        scores = split(q) @ split(k).transpose(-1, -2) / math.sqrt(width // heads)
        scores = scores.masked_fill(padding[:, None, None, :], float("-inf"))
        context = (drop(torch.softmax(scores, -1), f"attention.{index}") @ split(v)).transpose(1, 2).reshape(n_utt, frames, width)
        h = src + drop(context @ get(base + "self_attn.out_proj.weight").T + get(base + "self_attn.out_proj.bias"), f"attention_output.{index}")
        inner = activation(F.layer_norm(h, (width,), get(base + "norm2.weight"), get(base + "norm2.bias")) @ get(base + "linear1.weight").T + get(base + "linear1.bias"))
        inner = drop(inner, f"activation.{index}")
        h = h + drop(inner @ get(base + "linear2.weight").T + get(base + "linear2.bias"), f"feed_forward_output.{index}")
        outputs.append(F.layer_norm(h, (width,), get("_final_layer_norm.weight"), get("_final_layer_norm.bias")))
    return outputs, lengths
