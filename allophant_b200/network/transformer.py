"""From-scratch pre-LN transformer acoustic model (drop-in for ``allophant/network/acoustic_model.py:34-69, 552-759``).

``TransformerAcousticModel`` keeps the reference's module tree — ``nn.MultiheadAttention`` / ``nn.Linear`` /
``nn.LayerNorm`` instances created in the reference's order, so seeds, initialisation and ``state_dict`` keys match — but
only as parameter holders: the arithmetic is ``TransformerPlan``, a launch list over the same tcgen05 GEMM, flash
attention and row kernels the wav2vec2 path uses (``aph_gemm.cu``, ``aph_attention.cu``, ``aph_transformer.cu``).

Inference / frozen-encoder only in this build: the encoder has no CUDA backward yet, a forward in grad mode with
trainable encoder parameters raises ``NotImplementedError``; dropout layers are identities (``eval()`` arithmetic).
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

    training = False

    def __init__(self, model: "TransformerAcousticModel", n_utt: int, features: int, length: int, ldx: int, hidden_blocks: Dict[int, int]) -> None:
        self.model = model
        device = next(model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
        self.device = device
        self.n_utt, self.features, self.length, self.ldx = n_utt, features, length, ldx
        self.hidden_blocks = dict(hidden_blocks)
        self.generation = 0
        self.captured: Optional[List[Tensor]] = None
        if features != model._frontend_input_size:
            raise ValueError(f"expected {model._frontend_input_size} input features per frame, got {features}")
        # lengths through the (optional) sequential frontend: the time axis of the TENSOR follows the same formula
        seq, channels = length, model._frontend.output_dimensions
        self.stages: List[Dict[str, Any]] = []
        sequential = model._sequential_frontend
        if sequential is not None:
            for wrapper in sequential._layers.layers:
                module = wrapper.module
                if isinstance(module, Glu1d):
                    left, right = module.padding
                    kernel, stride = module.kernel_size, module.stride
                    conv = module._weights
                    out_channels = conv.out_channels // 2
                    out_len = (seq + left + right - kernel) // stride + 1
                    if (stride * channels) % 8 != 0 or (kernel * channels) % 8 != 0:
                        raise NotImplementedError("glu1d: stride * channels and kernel * channels have to be multiples of 8 (TMA alignment)")
                    self.stages.append(dict(kind="glu", module=module, left=left, right=right, kernel=kernel, stride=stride, in_len=seq,
                                            in_channels=channels, out_len=out_len, out_channels=out_channels))  # fmt: skip
                    seq, channels = out_len, out_channels
                elif isinstance(module, nn.Sequential):  # Transpose, LayerNorm, Transpose
                    self.stages.append(dict(kind="layer_norm", norm=module[1], channels=channels))
                elif isinstance(module, nn.Dropout):
                    continue  # identity in eval() arithmetic
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
        self._packed_version: Optional[Tuple[int, ...]] = None
        self._packed: Dict[str, Any] = {}

    # ------------------------------------------------------------------ weights
    def _pack(self) -> None:
        model = self.model
        version = tuple(p._version for p in model.parameters()) + (id(next(model.parameters())),)
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
            else:
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
    @torch.no_grad()
    def run(self, features: Tensor, lengths: Tensor, frames64: Tensor, capture: bool = False) -> None:
        model, N, d, M = self.model, self.n_utt, self.d, self.rows
        self._pack()
        packed = self._packed
        self.generation += 1
        self.captured = [] if capture else None
        eps = 1e-5
        lengths32 = lengths.to(torch.int32)
        ops.transpose_nfl(features, self.x_in, self.features)
        current, channels, seq = self.x_in, self.features, self.length
        frontend = model._frontend
        if isinstance(frontend, LinearFrontend):
            gamma, beta = packed["fe_ln"]
            normed = torch.empty(self.in_rows, channels, device=self.device, dtype=torch.bfloat16)
            ops.layernorm_any(current, channels, self.in_rows, channels, gamma, beta, frontend.layer_norm.eps, out_bf16=normed, ld_bf16=channels)
            neurons = frontend.output_dimensions
            out = torch.empty(self.in_rows, neurons, device=self.device, dtype=torch.float32)
            args = ops.make_gemm_args(normed, packed["fe_w"], a_rows=self.in_rows, a_inner=channels, a_row_stride=channels, bias=packed["fe_b"], out_f32=out, ld_f32=neurons)
            args.gelu = 3  # LeakyReLU(0.01)
            ops.run_gemm(args)
            current, channels = out, neurons
        for stage in self.stages:
            if stage["kind"] == "glu":
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
                lengths32 = torch.div(lengths32 + (left + right - kernel), stride, rounding_mode="floor") + 1
                current, channels, seq = out, out_channels, out_len
            else:
                gamma, beta = stage["affine"]
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
            g1, b1 = lw["ln1"]
            ops.layernorm_any(hidden, d, M, d, g1, b1, eps, out_f32=self.src, ld_f32=d, out_bf16=self.ln16, ld_bf16=d)
            ops.run_gemm(ops.make_qkv_args(self.ln16, lw["wqkv"], lw["bqkv"], self.q, self.k, self.v, rows=M, seq=seq, heads=self.heads))
            ops.attention(self.q, self.k, self.v, self.ctx, self.frames32, N, self.heads, seq)
            ops.run_gemm(ops.make_gemm_args(self.ctx, lw["wo"], a_rows=M, a_inner=d, a_row_stride=d, bias=lw["bo"], resid=self.src, ld_resid=d, out_f32=hidden, ld_f32=d))
            g2, b2 = lw["ln2"]
            ops.layernorm_any(hidden, d, M, d, g2, b2, eps, out_bf16=self.ln16, ld_bf16=d)
            args = ops.make_gemm_args(self.ln16, lw["w1"], a_rows=M, a_inner=d, a_row_stride=d, bias=lw["b1"], out_bf16=self.ffn, ld_bf16=self.ff)
            args.gelu = self.act
            ops.run_gemm(args)
            ops.run_gemm(ops.make_gemm_args(self.ffn, lw["w2"], a_rows=M, a_inner=self.ff, a_row_stride=self.ff, bias=lw["b2"], resid=hidden, ld_resid=d, out_f32=hidden, ld_f32=d))
            # acoustic_model.py:690: the final LayerNorm is applied to EVERY layer's output
            column = 0 if index == n_layers - 1 else self.hidden_blocks.get(index)
            if column is not None:
                ops.layernorm_any(hidden, d, M, d, final_gamma, final_beta, model._final_layer_norm.eps, out_bf16=self.x[:, column:], ld_bf16=self.ldx)
            if self.captured is not None:
                state = torch.empty(M, d, device=self.device, dtype=torch.float32)
                ops.layernorm_any(hidden, d, M, d, final_gamma, final_beta, model._final_layer_norm.eps, out_f32=state, ld_f32=d)
                self.captured.append(state)

    def backward(self, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError("the from-scratch transformer encoder has no CUDA backward pass in this build")


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

    def plan_for(self, n_utt: int, features: int, length: int, ldx: int, hidden_blocks: Dict[int, int]) -> TransformerPlan:
        key = (n_utt, features, length, ldx, tuple(sorted(hidden_blocks.items())), str(next(self.parameters()).device))
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = TransformerPlan(self, n_utt, features, length, ldx, hidden_blocks)
        return plan

    def encode(self, batch: Batch, ldx: int, hidden_blocks: Dict[int, int], capture: bool = False, training: bool = False, stochastic: Any = None) -> Tuple[TransformerPlan, Tensor]:
        features = batch.audio_features
        if not features.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the batch to the GPU (`batch.to('cuda')`)")
        if features.dim() != 3:
            raise ValueError(f"expected acoustic features of shape [batch, features, frames], got {tuple(features.shape)}")
        if training:
            raise NotImplementedError("the from-scratch transformer encoder has no CUDA backward pass in this build: freeze it or run under torch.no_grad()")
        features = features.float().contiguous()
        lengths = batch.lengths.to(device=features.device, dtype=torch.int64).contiguous()
        plan = self.plan_for(features.shape[0], features.shape[1], features.shape[2], ldx, hidden_blocks)
        frames = torch.empty(features.shape[0], device=features.device, dtype=torch.int64)
        plan.run(features, lengths, frames, capture)
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
