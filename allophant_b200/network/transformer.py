"""From-scratch pre-LN transformer acoustic model (drop-in for ``allophant/network/acoustic_model.py:34-69, 552-759``).

``TransformerAcousticModel`` keeps the reference's module tree — ``nn.MultiheadAttention`` / ``nn.Linear`` /
``nn.LayerNorm`` instances created in the reference's order, so seeds, initialisation and ``state_dict`` keys match — but
only as parameter holders: the arithmetic is ``TransformerPlan``, a launch list over the same tcgen05 GEMM, flash
attention and row kernels the wav2vec2 path uses (``aph_gemm.cu``, ``aph_attention.cu``, ``aph_transformer.cu``).

Training: ``TransformerPlan(training=True)`` keeps the activations and ``TransformerPlan.backward`` runs the backward
pass of the transformer layers, the final LayerNorm, the positional embeddings and the direct / linear front end on the same
GEMM (data-gradient / weight-gradient forms), flash-attention-backward and row kernels as the wav2vec2 path, and of the GLU
convolution stack (GLU', weight gradient with overlapping-row operands, data gradient + col2im with the reflections folded
back).  In ``train()`` mode the reference's dropout layers are applied with the counter-based masks of the wav2vec2 path
(GEMM-epilogue / attention / elementwise dropout; the backward pass regenerates them), in ``eval()`` mode they are identities.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import _lib, ops
from ..config import TransformerAcousticModelConfig
from ..dataset_processing import Batch
from .frontend import DirectFrontend, Frontend, Glu1d, LinearFrontend, SequentialFrontend, frontend_from_config


class SinusoidalPositionEmbeddings(nn.Module):
    """``acoustic_model.py:34-69`` (Vaswani et al. 2017): ``_bases[c] = exp(-(2 (c // 2)) ln(10000) / d)``."""

    _LOG_10000 = math.log(10000)

    def __init__(self, input_size: int):
        super().__init__()
        component = torch.exp(torch.arange(0, input_size, 2, dtype=torch.float) * -(self._LOG_10000 / input_size))
        self.register_buffer("_bases", torch.stack([component] * 2, 1).view(-1), persistent=False)
        self.embedding_size = input_size

    def get_positions(self, max_positions: int) -> Tensor:
        positions = torch.zeros(1, max_positions, self.embedding_size, device=self._bases.device)
        ops.add_sinusoidal(positions, self.embedding_size, 1, max_positions, self.embedding_size, self._bases)
        return positions.view(max_positions, -1)

    def forward(self, batch: Tensor) -> Tensor:
        return batch + self.get_positions(batch.size(0)).unsqueeze(1)


class PreLMTransformerEncoderLayer(nn.Module):
    """``acoustic_model.py:281-329``.  NOTE the reference's residual arithmetic: ``src = norm1(src)`` REPLACES the stream, so
    the attention branch is added to the NORMALISED input."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int = 2048, dropout: float = 0.1, activation: str = "relu", elementwise_affine: bool = False):
        super().__init__()
        if activation not in ("relu", "gelu"):
            raise RuntimeError("activation should be relu/gelu, not {}".format(activation))
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model, elementwise_affine=elementwise_affine)
        self.norm2 = nn.LayerNorm(d_model, elementwise_affine=elementwise_affine)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.activation_name = activation

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        raise RuntimeError("PreLMTransformerEncoderLayer is evaluated by TransformerAcousticModel (CUDA engine)")


class TransformerEncoderIntermediate(nn.Module):
    """``acoustic_model.py:332-353``: ``num_layers`` deep copies of one layer (``nn.TransformerEncoder`` semantics: all
    layers start from IDENTICAL parameters), every layer's output is returned."""

    def __init__(self, encoder_layer: PreLMTransformerEncoderLayer, num_layers: int):
        super().__init__()
        import copy

        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = None

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        raise RuntimeError("TransformerEncoderIntermediate is evaluated by TransformerAcousticModel (CUDA engine)")


def _affine(norm: nn.LayerNorm) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    if norm.weight is None:
        return None, None
    return norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous()


class TransformerPlan:
    """Workspaces + launch list of the transformer acoustic model for one ``(N, F, L)`` input shape.

    Exposes what the classifier runtime (``heads.py``) reads from a plan: ``x`` (bf16 ``[rows, ldx]``: final LayerNorm of the
    last layer in the first ``d`` columns, final LayerNorm of the kept layer outputs at ``hidden_blocks``), ``rows``,
    ``n_utt``, ``seq``, ``frames32``, ``generation``."""

    def __init__(self, model: "TransformerAcousticModel", n_utt: int, features: int, length: int, ldx: int, hidden_blocks: Dict[int, int],
                 training: bool = False) -> None:  # fmt: skip
        self.model = model
        self.training = training
        device = next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
        self.device = device
        self.n_utt, self.features, self.length, self.ldx = n_utt, features, length, ldx
        self.hidden_blocks = dict(hidden_blocks)
        self.generation = 0
        self.drop_seed: Optional[int] = None
        self.captured: Optional[List[Tensor]] = None
        if features != model._frontend_input_size:
            raise ValueError(f"expected {model._frontend_input_size} input features per frame, got {features}")
        # lengths through the (optional) sequential frontend: the time axis of the TENSOR follows the same formula
        seq, channels = length, model._frontend.output_dimensions
        self.stages: List[Dict[str, Any]] = []
        sequential = model._sequential_frontend
        if sequential is not None:
            for position, wrapper in enumerate(sequential._layers.layers):
                module = wrapper.module
                if isinstance(module, Glu1d):
                    left, right = module.padding
                    kernel, stride = module.kernel_size, module.stride
                    conv = module._weights
                    out_channels = conv.out_channels // 2
                    out_len = (seq + left + right - kernel) // stride + 1
                    if (stride * channels) % 8 != 0 or (kernel * channels) % 8 != 0:
                        raise NotImplementedError("glu1d: stride * channels and kernel * channels have to be multiples of 8 (TMA alignment)")
                    self.stages.append(dict(kind="glu", position=position, module=module, left=left, right=right, kernel=kernel, stride=stride, in_len=seq,
                                            in_channels=channels, out_len=out_len, out_channels=out_channels))  # fmt: skip
                    seq, channels = out_len, out_channels
                elif isinstance(module, nn.Sequential):  # Transpose, LayerNorm, Transpose
                    self.stages.append(dict(kind="layer_norm", position=position, norm=module[1], channels=channels))
                elif isinstance(module, nn.Dropout):
                    self.stages.append(dict(kind="dropout", position=position, rate=float(module.p)))  # identity unless train() mode
                else:
                    raise NotImplementedError(f"sequential frontend layer {type(module).__name__}")
        self.seq = seq
        self.rows = n_utt * seq
        d = channels
        self.d = d
        first = model._transformer.layers[0]
        heads = first.self_attn.num_heads
        if d != model.d_model:
            raise ValueError(f"the frontends produce {d} channels, the transformer expects {model.d_model}")
        if d % heads != 0 or d // heads != 64 or d % 256 != 0:
            raise NotImplementedError(
                f"the CUDA attention path supports a head dimension of 64 and model widths that are multiples of 256, got d_model {d} / {heads} heads"
            )
        self.heads = heads
        self.ff = first.linear1.out_features
        self.act = 1 if first.activation_name == "gelu" else 2
        bf16, f32 = torch.bfloat16, torch.float32
        z = lambda *shape, dtype=bf16: torch.zeros(*shape, device=device, dtype=dtype)  # noqa: E731
        M = self.rows
        self.frames32 = z(n_utt, dtype=torch.int32)
        self.in_rows = n_utt * length
        self.x_in = z(self.in_rows, features, dtype=f32)      # channels-last input features
        self.hidden = z(M, d, dtype=f32)                       # residual stream
        self.src = z(M, d, dtype=f32)                          # norm1 output (the stream the attention branch joins)
        self.ln16 = z(M, d)
        self.q, self.k, self.v = z(M * d), z(M * d), z(M * d)
        self.ctx = z(M, d)
        self.ffn = z(M, self.ff)
        self.x = z(M, ldx)
        if training:
            n_layers = len(model._transformer.layers)
            # everything the backward pass reads, per layer (only the attention probabilities are recomputed)
            self.saved = [
                dict(h_in=z(M, d, dtype=f32), src=z(M, d, dtype=f32), src16=z(M, d), q=z(M * d), k=z(M * d), v=z(M * d),
                     lse=z(n_utt * heads * seq, dtype=f32), ctx=z(M, d), h_mid=z(M, d, dtype=f32), ln2=z(M, d), pre=z(M, self.ff),
                     act=z(M, self.ff), h_out=z(M, d, dtype=f32))
                for _ in range(n_layers)
            ]  # fmt: skip
            self.fe_normed = z(self.in_rows, features) if isinstance(model._frontend, LinearFrontend) else None
            self.fe_out = z(self.in_rows, model._frontend.output_dimensions, dtype=f32) if isinstance(model._frontend, LinearFrontend) else None
            self.dh, self.dh16 = z(M, d, dtype=f32), z(M, d)
            self.d_ff, self.d_ln, self.d_ctx = z(M, self.ff), z(M, d, dtype=f32), z(M, d)
            self.dqkv, self.delta = z(M, 3 * d), z(n_utt * heads * seq, dtype=f32)
        self._packed_version: Optional[Tuple[int, ...]] = None
        self._packed: Dict[str, Any] = {}

    # ------------------------------------------------------------------ weights
    def _pack(self) -> None:
        model = self.model
        from ..engine import weight_generation  # fused optimiser steps write parameters behind torch's version counters

        version = tuple(p._version for p in model.parameters()) + (id(next(model.parameters())), weight_generation())
        if version == self._packed_version:
            return
        packed: Dict[str, Any] = {}
        cast = lambda w: w.detach().to(torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda w: w.detach().float().contiguous()  # noqa: E731
        frontend = model._frontend
        if isinstance(frontend, LinearFrontend):
            packed["fe_ln"] = _affine(frontend.layer_norm)
            packed["fe_w"], packed["fe_b"] = cast(frontend.linear.weight), f32(frontend.linear.bias)
        for index, stage in enumerate(self.stages):
            if stage["kind"] == "glu":
                conv = stage["module"]._weights
                stage["w"], stage["b"] = ops.pack_conv_weight(conv.weight), f32(conv.bias)
            elif stage["kind"] == "layer_norm":
                stage["affine"] = _affine(stage["norm"])
        layers = []
        for layer in model._transformer.layers:
            attention = layer.self_attn
            layers.append(
                dict(
                    ln1=_affine(layer.norm1), ln2=_affine(layer.norm2),
                    wqkv=cast(attention.in_proj_weight), bqkv=f32(attention.in_proj_bias),
                    wo=cast(attention.out_proj.weight), bo=f32(attention.out_proj.bias),
                    w1=cast(layer.linear1.weight), b1=f32(layer.linear1.bias), w2=cast(layer.linear2.weight), b2=f32(layer.linear2.bias),
                )
            )  # fmt: skip
        packed["layers"] = layers
        packed["final"] = _affine(model._final_layer_norm)
        self._packed, self._packed_version = packed, version

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _set_dropout(args: Any, drop: ops.Dropout) -> None:
        args.drop_threshold, args.drop_seed, args.drop_scale = drop.threshold, drop.seed, drop.scale

    SITE_FRONTEND_INPUT = 1_000_011
    SITE_MODEL_INPUT = 1_000_010
    SITE_SEQUENTIAL = 1_000_100  # + position of the Dropout layer in the sequential frontend

    def _site(self, rate: float, site: int) -> ops.Dropout:
        """The counter-based dropout of one site of the current train()-mode run (``aph_common.cuh``: ``drop_hash``)."""
        if self.drop_seed is None:
            return ops.NO_DROPOUT
        return ops.Dropout.site(rate, self.drop_seed, site)

    def _frontend_input_rate(self) -> float:
        frontend = self.model._frontend
        if isinstance(frontend, LinearFrontend):
            first = frontend._layer[0]
            return float(first.p) if isinstance(first, nn.Dropout) else 0.0
        return float(frontend._dropout.p) if getattr(frontend, "_dropout", None) is not None else 0.0

    @torch.no_grad()
    def run(self, features: Tensor, lengths: Tensor, frames64: Tensor, capture: bool = False, stochastic: Optional[int] = None) -> None:
        """``stochastic`` (training plans): the seed of a train()-mode run — the reference's dropout layers (front-end input
        dropouts, ``_input_dropout``, the Dropout layers of the sequential frontend, attention dropout and the three dropouts
        of every ``PreLMTransformerEncoderLayer``) are applied with counter-based masks the backward pass regenerates."""
        model, N, d, M = self.model, self.n_utt, self.d, self.rows
        self._pack()
        packed = self._packed
        self.generation += 1
        self.captured = [] if capture else None
        if stochastic is not None and not self.training:
            raise RuntimeError("train()-mode regularisation needs a training plan")
        self.drop_seed = stochastic
        layer_rate = float(model._transformer.layers[0].dropout.p)
        eps = 1e-5
        lengths32 = lengths.to(torch.int32)
        ops.transpose_nfl(features, self.x_in, self.features)
        current, channels, seq = self.x_in, self.features, self.length
        drop = self._site(self._frontend_input_rate(), self.SITE_FRONTEND_INPUT)
        if drop.threshold:
            ops.dropout_2d(current, channels, self.in_rows, channels, drop, out_f32=current, ld_f32=channels)
        frontend = model._frontend
        if isinstance(frontend, LinearFrontend):
            gamma, beta = packed["fe_ln"]
            normed = self.fe_normed if self.training else torch.empty(self.in_rows, channels, device=self.device, dtype=torch.bfloat16)
            ops.layernorm_any(current, channels, self.in_rows, channels, gamma, beta, frontend.layer_norm.eps, out_bf16=normed, ld_bf16=channels)
            neurons = frontend.output_dimensions
            out = self.fe_out if self.training else torch.empty(self.in_rows, neurons, device=self.device, dtype=torch.float32)
            args = ops.make_gemm_args(normed, packed["fe_w"], a_rows=self.in_rows, a_inner=channels, a_row_stride=channels, bias=packed["fe_b"], out_f32=out, ld_f32=neurons)
            args.gelu = 3  # LeakyReLU(0.01)
            ops.run_gemm(args)
            current, channels = out, neurons
        drop = self._site(float(model._input_dropout.p), self.SITE_MODEL_INPUT)  # acoustic_model.py:673
        if drop.threshold:
            ops.dropout_2d(current, channels, self.in_rows, channels, drop, out_f32=current, ld_f32=channels)
        for stage in self.stages:
            if stage["kind"] == "dropout":
                drop = self._site(stage["rate"], self.SITE_SEQUENTIAL + stage["position"])
                if drop.threshold:
                    ops.dropout_2d(current, channels, N * seq, channels, drop, out_f32=current, ld_f32=channels)
                stage["kept"] = dict(rows=N * seq, channels=channels)
            elif stage["kind"] == "glu":
                left, right, kernel, stride = stage["left"], stage["right"], stage["kernel"], stage["stride"]
                padded_len = seq + left + right
                padded = torch.empty(N * padded_len, channels, device=self.device, dtype=torch.bfloat16)
                ops.reflect_pad_bf16(current, channels, lengths32, N, seq, channels, left, right, stage["module"]._reflect_padding is not None, padded)
                out_len, out_channels = stage["out_len"], stage["out_channels"]
                gated = torch.empty(N * out_len, 2 * out_channels, device=self.device, dtype=torch.float32)
                ops.run_gemm(
                    ops.make_gemm_args(
                        padded, stage["w"], a_rows=out_len, a_inner=kernel * channels, a_row_stride=stride * channels, batch=N,
                        a_batch_stride=padded_len * channels, bias=stage["b"], out_f32=gated, ld_f32=2 * out_channels, out_batch_rows=out_len,
                    )
                )  # fmt: skip
                out = torch.empty(N * out_len, out_channels, device=self.device, dtype=torch.float32)
                ops.glu_rows(gated, 2 * out_channels, N * out_len, out_channels, out, out_channels)
                if self.training:  # the backward pass reads the padded input, the pre-gate output and the input lengths
                    stage["kept"] = dict(padded=padded, gated=gated, lengths32=lengths32, in_len=seq, in_channels=channels)
                lengths32 = torch.div(lengths32 + (left + right - kernel), stride, rounding_mode="floor") + 1
                current, channels, seq = out, out_channels, out_len
            else:
                gamma, beta = stage["affine"]
                if self.training:
                    stage["kept"] = dict(x=current, rows=N * seq)
                    normed = torch.empty_like(current)
                    ops.layernorm_any(current, channels, N * seq, channels, gamma, beta, stage["norm"].eps, out_f32=normed, ld_f32=channels)
                    current = normed
                else:
                    ops.layernorm_any(current, channels, N * seq, channels, gamma, beta, stage["norm"].eps, out_f32=current, ld_f32=channels)
        assert seq == self.seq and channels == d
        self.frames32.copy_(lengths32)
        frames64.copy_(lengths32)
        hidden = self.hidden
        hidden.copy_(current.view(M, d))
        if model._positional_embeddings is not None:
            ops.add_sinusoidal(hidden, d, N, seq, d, model._positional_embeddings._bases)
        final_gamma, final_beta = packed["final"]
        n_layers = len(packed["layers"])
        for index, lw in enumerate(packed["layers"]):
            if self.training:  # per-layer buffers: the backward pass reads them
                sv = self.saved[index]
                sv["h_in"].copy_(hidden)
                src, src16, q, k, v, ctx, ln2, ffn, pre, lse = sv["src"], sv["src16"], sv["q"], sv["k"], sv["v"], sv["ctx"], sv["ln2"], sv["act"], sv["pre"], sv["lse"]
                h_mid, h_out = sv["h_mid"], sv["h_out"]
            else:
                src, src16, q, k, v, ctx, ln2, ffn, pre, lse = self.src, self.ln16, self.q, self.k, self.v, self.ctx, self.ln16, self.ffn, None, None
                h_mid = h_out = hidden
            g1, b1 = lw["ln1"]
            ops.layernorm_any(hidden, d, M, d, g1, b1, eps, out_f32=src, ld_f32=d, out_bf16=src16, ld_bf16=d)
            ops.run_gemm(ops.make_qkv_args(src16, lw["wqkv"], lw["bqkv"], q, k, v, rows=M, seq=seq, heads=self.heads))
            ops.attention(q, k, v, ctx, self.frames32, N, self.heads, seq, lse, self._site(layer_rate, 8 * index))
            args = ops.make_gemm_args(ctx, lw["wo"], a_rows=M, a_inner=d, a_row_stride=d, bias=lw["bo"], resid=src, ld_resid=d, out_f32=h_mid, ld_f32=d)
            self._set_dropout(args, self._site(layer_rate, 8 * index + 1))  # dropout1
            ops.run_gemm(args)
            g2, b2 = lw["ln2"]
            ops.layernorm_any(h_mid, d, M, d, g2, b2, eps, out_bf16=ln2, ld_bf16=d)
            args = ops.make_gemm_args(ln2, lw["w1"], a_rows=M, a_inner=d, a_row_stride=d, bias=lw["b1"], out_bf16=ffn, ld_bf16=self.ff,
                                      aux_bf16=pre, ld_aux=self.ff)  # fmt: skip
            args.gelu = self.act
            self._set_dropout(args, self._site(layer_rate, 8 * index + 3))  # self.dropout, after the activation
            ops.run_gemm(args)
            args = ops.make_gemm_args(ffn, lw["w2"], a_rows=M, a_inner=self.ff, a_row_stride=self.ff, bias=lw["b2"], resid=h_mid, ld_resid=d, out_f32=h_out, ld_f32=d)
            self._set_dropout(args, self._site(layer_rate, 8 * index + 2))  # dropout2
            ops.run_gemm(args)
            if self.training:
                hidden = h_out
            # acoustic_model.py:690: the final LayerNorm is applied to EVERY layer's output
            column = 0 if index == n_layers - 1 else self.hidden_blocks.get(index)
            if column is not None:
                ops.layernorm_any(hidden, d, M, d, final_gamma, final_beta, model._final_layer_norm.eps, out_bf16=self.x[:, column:], ld_bf16=self.ldx)
            if self.captured is not None:
                state = torch.empty(M, d, device=self.device, dtype=torch.float32)
                ops.layernorm_any(hidden, d, M, d, final_gamma, final_beta, model._final_layer_norm.eps, out_f32=state, ld_f32=d)
                self.captured.append(state)

    @torch.no_grad()
    def backward(self, d_x: Tensor, need_encoder: bool = True, need_projection: bool = True, on_group_ready: Any = None) -> Dict[str, Tensor]:
        """Backward pass of ``run`` from ``d_x`` = dL/dX (fp32 ``[rows, ldx]``): fp32 gradients keyed by the parameter names of
        ``TransformerAcousticModel`` (``_transformer.layers.<i>.…``, ``_final_layer_norm.…``, ``_frontend._layer.<j>.…``).

        Layer arithmetic (``acoustic_model.py:313-329``): ``src = LN1(h); m = src + Wo attn(Wqkv src); out = m + W2 act(W1 LN2(m))`` —
        the attention branch joins the NORMALISED stream, so LN1's backward gets the sum of both paths and no residual
        by-passes it.  ``on_group_ready(flat, views)`` is called per layer like in ``EncoderPlan.backward``."""
        if not self.training:
            raise RuntimeError("this TransformerPlan was built for inference: no activations were kept")
        model, packed = self.model, self._packed
        N, M, d, FF, heads, seq = self.n_utt, self.rows, self.d, self.ff, self.heads, self.seq
        eps = 1e-5
        dev = d_x.device
        grads: Dict[str, Tensor] = {}
        layers = model._transformer.layers
        n_layers = len(layers)
        affine = model._final_layer_norm.weight is not None
        dh, dh16 = self.dh, self.dh16
        dh.zero_()

        def group(shapes):
            total = sum(int(torch.Size(shape).numel()) for _, shape in shapes)
            flat = torch.zeros(total, device=dev, dtype=torch.float32)
            views, offset = {}, 0
            for name, shape in shapes:
                count = int(torch.Size(shape).numel())
                views[name] = flat[offset : offset + count].view(shape)
                offset += count
            return flat, views

        def done(flat, views, prefix):
            named = {prefix + name: value for name, value in views.items()}
            grads.update(named)
            if on_group_ready is not None:
                on_group_ready(flat, named)

        def wgrad(out, dy, ld_dy, m, x, ld_x, n):
            ops.run_gemm(ops.make_wgrad_args(dy, x, out, rows=M, m=m, ld_dy=ld_dy, n=n, ld_x=ld_x, ld_out=n))

        def branch_gradient(drop: ops.Dropout) -> bool:
            """dh16 <- bf16(dh o keep * scale): gradient of a residual branch behind the forward's epilogue dropout."""
            if drop.threshold:
                ops.dropout_2d(dh, d, M, d, drop, out_bf16=dh16, ld_bf16=d)
                return True
            ops.cast_bf16_2d(dh, d, dh16, d, M, d)
            return False

        layer_rate = float(layers[0].dropout.p)
        final_flat, final_g = group([("weight", (d,)), ("bias", (d,))]) if affine else (None, {})
        final_gamma, _ = packed["final"]
        for index in reversed(range(n_layers)):
            lw, sv = packed["layers"][index], self.saved[index]
            shapes = [
                ("self_attn.in_proj_weight", (3 * d, d)), ("self_attn.in_proj_bias", (3 * d,)),
                ("self_attn.out_proj.weight", (d, d)), ("self_attn.out_proj.bias", (d,)),
                ("linear1.weight", (FF, d)), ("linear1.bias", (FF,)), ("linear2.weight", (d, FF)), ("linear2.bias", (d,)),
            ]  # fmt: skip
            if affine:
                shapes += [("norm1.weight", (d,)), ("norm1.bias", (d,)), ("norm2.weight", (d,)), ("norm2.bias", (d,))]
            flat, g = group(shapes)
            # every used layer output went through the final LayerNorm into a block of X (acoustic_model.py:690)
            column = 0 if index == n_layers - 1 else self.hidden_blocks.get(index)
            if column is not None:
                ops.layernorm_any_backward(sv["h_out"], d, d_x[:, column:], self.ldx, M, d, final_gamma, model._final_layer_norm.eps, dh, d, dh, d,
                                           final_g.get("weight"), final_g.get("bias"))  # fmt: skip
            # ---- feed forward: out = m + dropout2(W2 dropout(act(W1 LN2(m) + b1)) + b2)
            dropped = branch_gradient(self._site(layer_rate, 8 * index + 2))
            args = ops.make_dgrad_args(dh16, lw["w2"], rows=M, ld_dy=d, k=d, n=FF, ld_w=FF, gelu_bwd=sv["pre"], ld_gelu_bwd=FF, out_bf16=self.d_ff, ld_bf16=FF)
            args.act_bwd = self.act
            self._set_dropout(args, self._site(layer_rate, 8 * index + 3))  # the mask of the dropout behind the activation
            ops.run_gemm(args)
            wgrad(g["linear2.weight"], dh16, d, d, sv["act"], FF, FF)
            if dropped:
                ops.colsum_bf16(dh16, M, d, d, out=g["linear2.bias"])
            else:
                ops.colsum_f32(dh, M, d, d, out=g["linear2.bias"])
            wgrad(g["linear1.weight"], self.d_ff, FF, FF, sv["ln2"], d, d)
            ops.colsum_bf16(self.d_ff, M, FF, FF, out=g["linear1.bias"])
            ops.run_gemm(ops.make_dgrad_args(self.d_ff, lw["w1"], rows=M, ld_dy=FF, k=FF, n=d, ld_w=d, out_f32=self.d_ln, ld_f32=d))
            g2, _ = lw["ln2"]
            ops.layernorm_any_backward(sv["h_mid"], d, self.d_ln, d, M, d, g2, eps, dh, d, dh, d, g.get("norm2.weight"), g.get("norm2.bias"))
            # ---- attention block: m = src + dropout1(Wo attn(Wqkv src) + bo), src = LN1(h)
            dropped = branch_gradient(self._site(layer_rate, 8 * index + 1))
            ops.run_gemm(ops.make_dgrad_args(dh16, lw["wo"], rows=M, ld_dy=d, k=d, n=d, ld_w=d, out_bf16=self.d_ctx, ld_bf16=d))
            wgrad(g["self_attn.out_proj.weight"], dh16, d, d, sv["ctx"], d, d)
            if dropped:
                ops.colsum_bf16(dh16, M, d, d, out=g["self_attn.out_proj.bias"])
            else:
                ops.colsum_f32(dh, M, d, d, out=g["self_attn.out_proj.bias"])
            ops.attention_backward(sv["q"], sv["k"], sv["v"], sv["ctx"], self.d_ctx, sv["lse"], self.delta, self.dqkv, self.frames32, N, heads, seq,
                                   self._site(layer_rate, 8 * index))  # fmt: skip
            wgrad(g["self_attn.in_proj_weight"], self.dqkv, 3 * d, 3 * d, sv["src16"], d, d)
            ops.colsum_bf16(self.dqkv, M, 3 * d, 3 * d, out=g["self_attn.in_proj_bias"])
            # d(src) = dh (the stream the branch joined) + dqkv Wqkv, in one GEMM with the residual epilogue
            ops.run_gemm(ops.make_dgrad_args(self.dqkv, lw["wqkv"], rows=M, ld_dy=3 * d, k=3 * d, n=d, ld_w=d, resid=dh, ld_resid=d, out_f32=self.d_ln, ld_f32=d))
            g1, _ = lw["ln1"]
            ops.layernorm_any_backward(sv["h_in"], d, self.d_ln, d, M, d, g1, eps, None, 0, dh, d, g.get("norm1.weight"), g.get("norm1.bias"))
            done(flat, g, f"_transformer.layers.{index}.")
        if affine:
            done(final_flat, final_g, "_final_layer_norm.")
        # positional embeddings are an additive constant; dh is now the gradient of the last front-end stage's output
        d_cur, cur_rows, cur_channels = dh, M, d
        for stage in reversed(self.stages):
            kept, position = stage["kept"], stage["position"]
            if stage["kind"] == "dropout":
                drop = self._site(stage["rate"], self.SITE_SEQUENTIAL + position)
                if drop.threshold:
                    ops.dropout_2d(d_cur, cur_channels, cur_rows, cur_channels, drop, out_f32=d_cur, ld_f32=cur_channels)
                continue
            if stage["kind"] == "layer_norm":
                norm = stage["norm"]
                gamma, _ = stage["affine"]
                flat, g = group([("weight", (cur_channels,)), ("bias", (cur_channels,))]) if gamma is not None else (None, {})
                d_in = torch.empty(cur_rows, cur_channels, device=dev, dtype=torch.float32)
                ops.layernorm_any_backward(kept["x"], cur_channels, d_cur, cur_channels, cur_rows, cur_channels, gamma, norm.eps, None, 0, d_in, cur_channels,
                                           g.get("weight"), g.get("bias"))  # fmt: skip
                if gamma is not None:
                    done(flat, g, f"_sequential_frontend._layers.layers.{position}.module.1.")
                d_cur = d_in
                continue
            module = stage["module"]
            left, right, kernel, stride = stage["left"], stage["right"], stage["kernel"], stage["stride"]
            in_len, in_channels, out_len, out_channels = kept["in_len"], kept["in_channels"], stage["out_len"], stage["out_channels"]
            rows_out = N * out_len
            width = kernel * in_channels
            d_gated = torch.empty(rows_out, 2 * out_channels, device=dev, dtype=torch.bfloat16)
            ops.glu_backward_bf16(kept["gated"], 2 * out_channels, d_cur, cur_channels, rows_out, out_channels, d_gated, 2 * out_channels)
            flat, g = group([("weight", (2 * out_channels, kernel, in_channels)), ("bias", (2 * out_channels,))])
            ops.colsum_bf16(d_gated, rows_out, 2 * out_channels, 2 * out_channels, out=g["bias"])
            # dW[o][j][c] = sum_(n,t) dY[n][t][o] * Xpad[n][t*stride + j][c]: both operands frame-major, one K segment per utterance,
            # the windows of Xpad are overlapping rows of stride `stride * C` (the same tensor map trick as the forward conv)
            args = ops.make_wgrad_args(d_gated, kept["padded"], g["weight"].view(2 * out_channels, width), rows=out_len, m=2 * out_channels,
                                       ld_dy=2 * out_channels, n=width, ld_x=stride * in_channels, ld_out=width)  # fmt: skip
            args.k_batch, args.a_batch_stride, args.b_seg_stride = N, out_len * 2 * out_channels, (in_len + left + right) * in_channels
            ops.run_gemm(args)
            weight_gradient = g.pop("weight").permute(0, 2, 1).contiguous()  # [2O, k, C] -> Conv1d's [2O, C, k]
            named = {f"_sequential_frontend._layers.layers.{position}.module._weights.weight": weight_gradient,
                     f"_sequential_frontend._layers.layers.{position}.module._weights.bias": g["bias"]}  # fmt: skip
            grads.update(named)
            if on_group_ready is not None:
                on_group_ready(weight_gradient.view(-1), {k: v for k, v in named.items() if k.endswith("weight")})
                on_group_ready(g["bias"], {k: v for k, v in named.items() if k.endswith("bias")})
            first_stage = all(earlier["kind"] == "dropout" for earlier in self.stages[: self.stages.index(stage)])
            frontend_trains = isinstance(model._frontend, LinearFrontend) and any(p.requires_grad for p in model._frontend.parameters())
            if first_stage and not frontend_trains:
                d_cur = None  # nothing upstream needs a gradient
                break
            d_cols = torch.empty(rows_out, width, device=dev, dtype=torch.float32)
            ops.run_gemm(ops.make_dgrad_args(d_gated, stage["w"], rows=rows_out, ld_dy=2 * out_channels, k=2 * out_channels, n=width, ld_w=width, out_f32=d_cols, ld_f32=width))
            d_in = torch.empty(N * in_len, in_channels, device=dev, dtype=torch.float32)
            ops.conv_input_backward(d_cols, kept["lengths32"], N, in_len, in_channels, out_len, kernel, stride, left, right, module._reflect_padding is not None, d_in, in_channels)
            d_cur, cur_rows, cur_channels = d_in, N * in_len, in_channels
        frontend = model._frontend
        if d_cur is not None and isinstance(frontend, LinearFrontend) and any(p.requires_grad for p in frontend.parameters()):
            neurons, features = frontend.output_dimensions, self.features
            rows_in = self.in_rows
            drop = self._site(float(model._input_dropout.p), self.SITE_MODEL_INPUT)
            if drop.threshold:
                ops.dropout_2d(d_cur, neurons, rows_in, neurons, drop, out_f32=d_cur, ld_f32=neurons)
            names = {id(module): index for index, module in enumerate(frontend._layer)}
            linear_index, norm_index = names[id(frontend.linear)], names[id(frontend.layer_norm)]
            shapes = [(f"{linear_index}.weight", (neurons, features)), (f"{linear_index}.bias", (neurons,))]
            norm_affine = frontend.layer_norm.weight is not None
            if norm_affine:
                shapes += [(f"{norm_index}.weight", (features,)), (f"{norm_index}.bias", (features,))]
            flat, g = group(shapes)
            d16 = torch.empty(rows_in, neurons, device=dev, dtype=torch.bfloat16)
            ops.activation_backward(d_cur, neurons, self.fe_out, neurons, rows_in, neurons, 3, d16, neurons)  # LeakyReLU, decided from its output
            ops.run_gemm(ops.make_wgrad_args(d16, self.fe_normed, g[f"{linear_index}.weight"], rows=rows_in, m=neurons, ld_dy=neurons, n=features, ld_x=features, ld_out=features))
            ops.colsum_f32(d_cur, rows_in, neurons, neurons, out=g[f"{linear_index}.bias"])
            if norm_affine:
                d_normed = torch.empty(rows_in, features, device=dev, dtype=torch.float32)
                ops.run_gemm(ops.make_dgrad_args(d16, packed["fe_w"], rows=rows_in, ld_dy=neurons, k=neurons, n=features, ld_w=features, out_f32=d_normed, ld_f32=features))
                gamma, _ = packed["fe_ln"]
                ops.layernorm_any_backward(self.x_in, features, d_normed, features, rows_in, features, gamma, frontend.layer_norm.eps, None, 0, None, 0,
                                           g[f"{norm_index}.weight"], g[f"{norm_index}.bias"])  # fmt: skip
            done(flat, g, "_frontend._layer.")
        return grads


class TransformerAcousticModel(nn.Module):
    """``acoustic_model.py:643-759``."""

    def __init__(
        self,
        frontend: Frontend,
        transformer: TransformerEncoderIntermediate,
        sequential_frontend: Optional[SequentialFrontend] = None,
        input_dropout_rate: float = 0,
        use_positional_embeddings: bool = True,
        elementwise_affine: bool = False,
        feature_size: Optional[int] = None,
    ) -> None:
        super().__init__()
        self._input_dropout = nn.Dropout(input_dropout_rate)
        self._frontend = frontend
        self._transformer = transformer
        self._feature_size = frontend.output_dimensions
        model_width = self._feature_size if sequential_frontend is None else sequential_frontend.output_dimensions
        self._final_layer_norm = nn.LayerNorm(model_width, elementwise_affine=elementwise_affine)
        self._positional_embeddings = SinusoidalPositionEmbeddings(model_width) if use_positional_embeddings else None
        self._sequential_frontend = sequential_frontend
        self._upscale_factor = 1 if sequential_frontend is None else sequential_frontend.upscale_factor
        self._d_model = transformer.layers[0].linear1.in_features
        self._output_size = transformer.layers[-1].linear2.out_features
        if isinstance(frontend, LinearFrontend):
            self._frontend_input_size = frontend.linear.in_features
        else:
            self._frontend_input_size = frontend.output_dimensions if feature_size is None else feature_size
        self._plans: Dict[Tuple[Any, ...], TransformerPlan] = {}

    # reference properties (acoustic_model.py:620-640, 664-670)
    @property
    def feature_size(self) -> int:
        return self._feature_size

    @property
    def output_size(self) -> int:
        return self._output_size

    @property
    def d_model(self) -> int:
        return self._d_model

    @property
    def upscale_factor(self) -> float:
        return self._upscale_factor

    @property
    def hidden_state_count(self) -> int:
        """Entries of the list ``forward`` returns: one per transformer layer (no embedding entry, unlike wav2vec2)."""
        return len(self._transformer.layers)

    def encoder_parameters_require_grad(self) -> bool:
        return any(p.requires_grad for p in self.parameters())

    def downsampled_lengths(self, lengths: Tensor) -> Tensor:
        lengths = self._frontend.lengths(lengths)
        if self._sequential_frontend is None:
            return lengths
        return self._sequential_frontend.downsampled_lengths(lengths)

    def plan_for(self, n_utt: int, features: int, length: int, ldx: int, hidden_blocks: Dict[int, int], training: bool = False) -> TransformerPlan:
        key = (n_utt, features, length, ldx, tuple(sorted(hidden_blocks.items())), str(next(self.parameters()).device), training)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = TransformerPlan(self, n_utt, features, length, ldx, hidden_blocks, training)
        return plan

    def encode(self, batch: Batch, ldx: int, hidden_blocks: Dict[int, int], capture: bool = False, training: bool = False, stochastic: Any = None) -> Tuple[TransformerPlan, Tensor]:
        features = batch.audio_features
        if not features.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the batch to the GPU (`batch.to('cuda')`)")
        if features.dim() != 3:
            raise ValueError(f"expected acoustic features of shape [batch, features, frames], got {tuple(features.shape)}")
        features = features.float().contiguous()
        lengths = batch.lengths.to(device=features.device, dtype=torch.int64).contiguous()
        plan = self.plan_for(features.shape[0], features.shape[1], features.shape[2], ldx, hidden_blocks, training)
        frames = torch.empty(features.shape[0], device=features.device, dtype=torch.int64)
        plan.run(features, lengths, frames, capture, stochastic)
        return plan, frames

    def forward(self, batch: Batch, _predict: bool = False) -> Tuple[List[Tensor], Tensor]:
        """Every layer's output after the final LayerNorm, time-first ``[L', N, d]``, and the frame counts
        (``acoustic_model.py:669-691``)."""
        plan, frames = self.encode(batch, self._d_model, {}, capture=True)
        assert plan.captured is not None
        return [state.view(plan.n_utt, plan.seq, -1).transpose(0, 1) for state in plan.captured], frames

    @classmethod
    def from_config(cls, layer_config: TransformerAcousticModelConfig, feature_size: int) -> "TransformerAcousticModel":
        transformer_config = layer_config.transformer
        frontend = frontend_from_config(layer_config.frontend, feature_size, layer_config.elementwise_affine)
        previous_output_size = frontend.output_dimensions
        if layer_config.sequential_frontend is not None:
            sequential_frontend = SequentialFrontend.from_config(layer_config.sequential_frontend, previous_output_size)
            previous_output_size = sequential_frontend.output_dimensions
        else:
            sequential_frontend = None
        return cls(
            frontend,
            TransformerEncoderIntermediate(
                PreLMTransformerEncoderLayer(
                    previous_output_size,
                    transformer_config.heads,
                    transformer_config.feedforward_neurons,
                    transformer_config.dropout_rate,
                    transformer_config.activation,
                    layer_config.elementwise_affine,
                ),
                transformer_config.num_layers,
            ),
            sequential_frontend,
            transformer_config.dropout_rate,
            transformer_config.positional_embeddings,
            layer_config.elementwise_affine,
            feature_size,
        )
