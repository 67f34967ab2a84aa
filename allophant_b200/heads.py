"""Runtime of the classifier heads (``HierarchicalProjection.forward``, ``acoustic_model.py:471-524``).

The reference evaluates one ``nn.Linear`` per classifier (37 launches of N≈4 columns).  Here the
classifiers are grouped into dependency levels and every level is ONE tcgen05 GEMM over a shared
bf16 feature matrix

    X = [ OUTPUT (final LayerNorm) | OUTPUT_i blocks | softmax(dependency logits) blocks | 0-pad ]

whose per-classifier weight rows are scattered into the columns of their dependencies (zeros
elsewhere), so ``cat([...], -1) @ W.T`` never materialises.  Composed phoneme logits
(``EmbeddingCompositionLayer``, 219-234) are a second GEMM against the gather-summed embedding
table; ``log_softmax`` of all heads is one launch that also emits the per-frame argmax used
by greedy CTC decoding.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import ops
from .config import ProjectionEntryConfig
from .dataset_processing import Batch

_OUTPUT = ProjectionEntryConfig.OUTPUT_DEPENDENCY
_PATTERN = ProjectionEntryConfig.OUTPUT_PATTERN


def _version(tensor: Tensor) -> int:
    """Version counter, or -1 for inference tensors (which do not track one)."""
    try:
        return tensor._version
    except RuntimeError:
        return -1


def _round_up(value: int, multiple: int) -> int:
    return (value + multiple - 1) // multiple * multiple


@dataclass
class _LevelLayout:
    specs: List[Any]
    offsets: Dict[str, int]  # classifier name -> first column in the level's logits
    n_pad: int
    has_composition: bool
    feeds_later: List[str] = field(default_factory=list)  # classifiers of this level used as dependencies


class _ProjectionFunction(torch.autograd.Function):
    """All classifier heads as one autograd node: forward and backward run in ``liballophant_b200.so``."""

    @staticmethod
    def forward(ctx, runtime, state, *parameters):
        ctx.runtime = runtime
        ctx.state = state
        outputs = runtime._differentiable_forward_impl(state)
        return outputs

    @staticmethod
    def backward(ctx, *grads):
        parameter_grads = ctx.runtime._differentiable_backward_impl(ctx.state, grads)
        return (None, None, *parameter_grads)


class HeadsRuntime:
    def __init__(self, model: Any) -> None:
        self.model = model
        self._layout_ready = False
        self._weights_version: Optional[Tuple[int, ...]] = None
        self._composed_cache: Dict[Tuple[Any, ...], Tuple[Tensor, int]] = {}
        self._checked_inventories: set = set()  # index tensors whose categories were range-checked already
        self._index_cache: Dict[Tuple[Any, ...], Dict[str, Tensor]] = {}
        self.gradient_reducer: Any = None  # allophant_b200.distributed.GradientReducer (data-parallel training)
        self.skip_layers_override: Optional[Sequence[bool]] = None  # explicit LayerDrop decisions for train() mode (tests)
        self.last_regularisation: Dict[str, Any] = {}

    # ------------------------------------------------------------------ static layout
    def _build_layout(self) -> None:
        projection = self.model._projection
        hidden = projection._output_features
        n_layers = self.model._acoustic_model.hidden_state_count - 1  # index of the LAST hidden state (= OUTPUT)
        column = hidden
        self.x_cols: Dict[str, int] = {_OUTPUT: 0}
        self.hidden_blocks: Dict[int, int] = {}
        for name in projection._output_dependencies:
            if name == _OUTPUT:
                continue
            index = int(_PATTERN.match(name).group(1))
            if index > n_layers:
                raise ValueError(f"{name} does not exist: the encoder has {n_layers + 1} hidden states")
            if index == n_layers:
                self.x_cols[name] = 0  # the last hidden state IS the final LayerNorm output
            else:
                self.x_cols[name] = column
                self.hidden_blocks[index] = column
                column += hidden
        self.dep_cols: Dict[str, Tuple[int, int]] = {}
        for spec in projection._specs:
            for dependency in spec.dependencies:
                if _PATTERN.match(dependency.name) or dependency.name in self.dep_cols:
                    continue
                self.dep_cols[dependency.name] = (column, dependency.size)
                column += dependency.size
        self.ldx = _round_up(column, 64)

        by_level: Dict[int, List[Any]] = {}
        for spec in projection._specs:
            by_level.setdefault(spec.level, []).append(spec)
        self.levels: List[_LevelLayout] = []
        for level in sorted(by_level):
            specs = by_level[level]
            # composition projections first: their bf16 output is a GEMM A operand and must start 16-byte aligned
            ordered = sorted(specs, key=lambda s: 0 if projection._layers[s.name]._composition_layer is not None else 1)
            offsets, column = {}, 0
            for spec in ordered:
                if projection._layers[spec.name]._composition_layer is not None:
                    column = _round_up(column, 8)
                offsets[spec.name] = column
                column += spec.out_features
            layout = _LevelLayout(
                ordered,
                offsets,
                _round_up(column, 8),
                any(projection._layers[s.name]._composition_layer is not None for s in ordered),
            )
            layout.feeds_later = [s.name for s in ordered if s.name in self.dep_cols]
            self.levels.append(layout)
        self._layout_ready = True

    # ------------------------------------------------------------------ weights
    def _params_version(self) -> Tuple[int, ...]:
        params = getattr(self, "_projection_params", None)
        if params is None:
            params = self._projection_params = list(self.model._projection.parameters())
        from .engine import weight_generation

        return (weight_generation(),) + tuple(_version(p) for p in params) + tuple(p.data_ptr() for p in params)

    @torch.no_grad()
    def _ensure_weights(self, device: torch.device) -> None:
        version = self._params_version()
        if version == self._weights_version:
            return
        projection = self.model._projection
        self.level_w: List[Tensor] = []
        self.level_b: List[Tensor] = []
        for layout in self.levels:
            weight = torch.zeros(layout.n_pad, self.ldx, device=device, dtype=torch.float32)
            bias = torch.zeros(layout.n_pad, device=device, dtype=torch.float32)
            # the classifiers' weight / bias blocks join the level matrices in two launches (a training step rebuilds them after
            # every optimizer step: one strided add and one copy per classifier were 74 launches on its critical path)
            w_src, w_dst, b_src, b_dst = [], [], [], []
            seen_targets: set = set()
            late_adds: List[Tuple[Tensor, Tensor]] = []
            for spec in layout.specs:
                linear = projection._layers[spec.name]._time_distributed_layer
                if not isinstance(linear, torch.nn.Linear):
                    # time layer (ProjectingMultiheadAttention): its input projection is the Linear of the level GEMM, the
                    # attention behind it runs on the level's columns afterwards (_run_time_layer)
                    self._pack_time_layer(spec.name, linear, device)
                    linear = linear.input_projection
                if not linear.weight.is_cuda:
                    raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
                row = layout.offsets[spec.name]
                source_column = 0
                for dependency in spec.dependencies:
                    if _PATTERN.match(dependency.name):
                        target = self.x_cols[dependency.name]
                    else:
                        target = self.dep_cols[dependency.name][0]
                    source = linear.weight.detach()
                    # (two dependencies on the same columns, e.g. the output and the last hidden state, must not race in one launch)
                    first_use = (row, target) not in seen_targets
                    seen_targets.add((row, target))
                    if first_use and source.dtype == torch.float32 and source.is_contiguous():
                        w_src.append((source, source.shape[1], source_column, dependency.size, spec.out_features))
                        w_dst.append((weight, self.ldx, row * self.ldx + target, dependency.size, spec.out_features))
                    else:
                        late_adds.append((weight[row : row + spec.out_features, target : target + dependency.size],
                                          source[:, source_column : source_column + dependency.size].float()))  # fmt: skip
                    source_column += dependency.size
                source_bias = linear.bias.detach()
                if source_bias.dtype == torch.float32 and source_bias.is_contiguous():
                    b_src.append((source_bias, spec.out_features, 0, spec.out_features, 1))
                    b_dst.append((bias, layout.n_pad, row, spec.out_features, 1))
                else:
                    bias[row : row + spec.out_features] = source_bias.float()
            ops.copy_head_blocks(w_src, w_dst, 0, accumulate=True)
            ops.copy_head_blocks(b_src, b_dst, 0)
            for destination, addend in late_adds:
                destination += addend
            self.level_w.append(ops.cast_bf16(weight))
            self.level_b.append(bias)
        self._composed_cache.clear()
        self._weights_version = version

    # ------------------------------------------------------------------ time layers (acoustic_model.py:237-268)
    def _pack_time_layer(self, name: str, layer: Any, device: torch.device) -> None:
        """bf16 operands of nn.MultiheadAttention's in_proj / out_proj, zero-padded to the GEMM's granularity (8)."""
        hidden = layer.hidden_dimensions
        h_pad, qkv_pad = _round_up(hidden, 8), _round_up(3 * hidden, 8)
        attention = layer.attention
        in_w = torch.zeros(qkv_pad, h_pad, device=device, dtype=torch.float32)
        in_w[: 3 * hidden, :hidden] = attention.in_proj_weight.detach().float()
        in_b = torch.zeros(qkv_pad, device=device, dtype=torch.float32)
        in_b[: 3 * hidden] = attention.in_proj_bias.detach().float()
        out_w = torch.zeros(h_pad, h_pad, device=device, dtype=torch.float32)
        out_w[:hidden, :hidden] = attention.out_proj.weight.detach().float()
        out_b = torch.zeros(h_pad, device=device, dtype=torch.float32)
        out_b[:hidden] = attention.out_proj.bias.detach().float()
        if not hasattr(self, "_time_layers"):
            self._time_layers: Dict[str, Dict[str, Any]] = {}
        self._time_layers[name] = dict(
            hidden=hidden, h_pad=h_pad, qkv_pad=qkv_pad, heads=layer.num_heads, in_w=ops.cast_bf16(in_w), in_b=in_b, out_w=ops.cast_bf16(out_w), out_b=out_b,
            gamma=layer.layer_norm.weight.detach().float().contiguous(), beta=layer.layer_norm.bias.detach().float().contiguous(),
            eps=layer.layer_norm.eps, bases=None if layer.positional_embeddings is None else layer.positional_embeddings._bases.to(device),
        )  # fmt: skip

    def _run_time_layer(self, name: str, plan, logits: Tensor, ld: int, offset: int, logits_bf16: Optional[Tensor]) -> None:
        """``logits[:, offset : offset + hidden]`` holds the input projection of head ``name``; on return it holds
        ``MultiheadAttention(LayerNorm(projection) (+ positions))`` over the frames of every utterance (eval(): no dropout)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model._projection._layers[name].parameters()):
            raise NotImplementedError("training through a multi-head-attention time layer is not built (inference only)")
        t = self._time_layers[name]
        rows, n_utt, seq, hidden, h_pad = plan.rows, plan.n_utt, plan.seq, t["hidden"], t["h_pad"]
        device = logits.device
        projected = logits[:, offset:]
        normalised = torch.zeros(rows, h_pad, device=device, dtype=torch.float32) if t["bases"] is not None else None
        operand = torch.zeros(rows, h_pad, device=device, dtype=torch.bfloat16)
        if t["bases"] is None:
            ops.layernorm_any(projected, ld, rows, hidden, t["gamma"], t["beta"], t["eps"], out_bf16=operand, ld_bf16=h_pad)
        else:
            ops.layernorm_any(projected, ld, rows, hidden, t["gamma"], t["beta"], t["eps"], out_f32=normalised, ld_f32=h_pad)
            ops.add_sinusoidal(normalised, h_pad, n_utt, seq, hidden, t["bases"])
            ops.cast_bf16_2d(normalised, h_pad, operand, h_pad, rows, h_pad)
        qkv = torch.empty(rows, t["qkv_pad"], device=device, dtype=torch.float32)
        ops.run_gemm(ops.make_gemm_args(operand, t["in_w"], a_rows=rows, a_inner=h_pad, a_row_stride=h_pad, bias=t["in_b"], out_f32=qkv, ld_f32=t["qkv_pad"]))
        context = torch.zeros(rows, h_pad, device=device, dtype=torch.bfloat16)
        ops.attention_small(qkv, t["qkv_pad"], context, h_pad, plan.frames32, n_utt, t["heads"], seq, hidden // t["heads"])
        attended = torch.empty(rows, h_pad, device=device, dtype=torch.float32)
        ops.run_gemm(ops.make_gemm_args(context, t["out_w"], a_rows=rows, a_inner=h_pad, a_row_stride=h_pad, bias=t["out_b"], out_f32=attended, ld_f32=h_pad))
        projected[:, :hidden].copy_(attended[:, :hidden])
        if logits_bf16 is not None:
            logits_bf16[:, offset : offset + hidden].copy_(attended[:, :hidden])

    def _composed_embeddings(self, name: str, target_feature_indices: Optional[Tensor], device: torch.device) -> Tuple[Tensor, int]:
        """bf16 [Vpad, E] table of blank + composed phoneme embeddings and the number of classes V+1."""
        layer = self.model._projection._layers[name]._composition_layer
        weight = layer._attribute_embeddings.weight
        if target_feature_indices is None:
            indices, offsets = layer._dense_feature_table, None
        else:
            indices, offsets = target_feature_indices, layer._category_offsets
        key = (name, indices.data_ptr(), _version(indices), tuple(indices.shape), offsets is None, _version(weight), weight.data_ptr())
        cached = self._composed_cache.get(key)
        if cached is not None:
            return cached
        indices_dev = indices.to(device=device, dtype=torch.int64).contiguous()
        if indices_dev.dim() != 2 or indices_dev.shape[1] != layer._category_offsets.shape[1]:
            raise ValueError(
                f"target_feature_indices must have shape [phonemes, {layer._category_offsets.shape[1]}], got {tuple(indices_dev.shape)}"
            )
        offsets_dev = None if offsets is None else offsets.to(device=device, dtype=torch.int64).contiguous().view(-1)
        classes = indices_dev.shape[0] + 1
        rows = _round_up(classes, 8)
        table = torch.empty(rows, weight.shape[1], device=device, dtype=torch.bfloat16)
        err = torch.zeros(1, device=device, dtype=torch.int32)
        ops.compose_embeddings(weight.detach().float().contiguous(), indices_dev, offsets_dev, rows, err, out_bf16=table)
        # one-time check per inventory (EmbeddingBag raises on out-of-range indices too).  Only the first composition of an
        # inventory reads the flag back: while training the table is recomposed after every optimizer step, and a
        # device-to-host read there would drain the stream once per step
        checked = key[1:5] + (int(weight.shape[0]),)
        if checked not in self._checked_inventories:
            if int(err.item()) != 0:
                raise IndexError("target_feature_indices contain a category outside the attribute embedding table")
            if len(self._checked_inventories) >= 64:
                self._checked_inventories.clear()
            self._checked_inventories.add(checked)
        if len(self._composed_cache) >= 8:
            self._composed_cache.pop(next(iter(self._composed_cache)))
        self._composed_cache[key] = (table, classes)
        return table, classes

    def _int_tensor(self, key: Tuple[Any, ...], values: List[int], device: torch.device, dtype: torch.dtype) -> Tensor:
        cache = self._index_cache.setdefault(("t",) + key, {})
        tensor = cache.get("v")
        if tensor is None:
            tensor = torch.tensor(values, device=device, dtype=dtype)
            cache["v"] = tensor
        return tensor

    # ------------------------------------------------------------------ level GEMMs
    def _run_levels(self, plan, batch: Batch, target_feature_indices: Optional[Tensor], predict: bool, keep: Optional[Dict[str, Any]]):
        """Evaluates every classifier level.  Returns the heads list [(name, buffer, ld, column, width)] in the
        reference's output order; with ``keep`` (a dict) also records what the backward pass needs."""
        projection = self.model._projection
        device = plan.x.device
        rows, n_utt, seq = plan.rows, plan.n_utt, plan.seq
        skip = 0 if projection._dependency_blanks else projection._blank_offset
        # (name, logits buffer, leading dim, first column, width) in topological order
        heads: List[Tuple[str, Tensor, int, int, int]] = []
        level_logits: List[Tensor] = []
        level_bf16: List[Optional[Tensor]] = []
        composed_info: Dict[str, Dict[str, Any]] = {}
        produced_all: Dict[str, Tuple[Tensor, int, int, int]] = {}
        for level_index, layout in enumerate(self.levels):
            logits = torch.empty(rows, layout.n_pad, device=device, dtype=torch.float32)
            logits_bf16 = torch.empty(rows, layout.n_pad, device=device, dtype=torch.bfloat16) if layout.has_composition else None
            args = ops.make_gemm_args(
                plan.x,
                self.level_w[level_index],
                a_rows=rows,
                a_inner=self.ldx,
                a_row_stride=self.ldx,
                bias=self.level_b[level_index],
                out_f32=logits,
                ld_f32=layout.n_pad,
                out_bf16=logits_bf16,
                ld_bf16=layout.n_pad,
            )
            ops.run_gemm(args)
            level_logits.append(logits)
            level_bf16.append(logits_bf16)
            produced: Dict[str, Tuple[Tensor, int, int, int]] = {}
            for spec in layout.specs:
                classifier = projection._layers[spec.name]
                offset = layout.offsets[spec.name]
                if classifier._lengths_required:  # time layer: attention over the frames on this head's columns
                    if keep is not None:
                        raise NotImplementedError("training through a multi-head-attention time layer is not built (inference only)")
                    self._run_time_layer(spec.name, plan, logits, layout.n_pad, offset, logits_bf16)
                if classifier._composition_layer is not None:
                    table, classes = self._composed_embeddings(spec.name, target_feature_indices, device)
                    embedding = classifier._composition_layer.embedding_size
                    composed = torch.empty(rows, table.shape[0], device=device, dtype=torch.float32)
                    assert logits_bf16 is not None
                    comp_args = ops.make_gemm_args(
                        logits_bf16[:, offset:],
                        table,
                        a_rows=rows,
                        a_inner=embedding,
                        a_row_stride=layout.n_pad,
                        scale=1.0 / math.sqrt(embedding),
                        out_f32=composed,
                        ld_f32=table.shape[0],
                    )
                    ops.run_gemm(comp_args)
                    composed_info[spec.name] = dict(table=table, classes=classes, logits=composed, level=level_index, offset=offset)
                    produced[spec.name] = (composed, table.shape[0], 0, classes)
                else:
                    produced[spec.name] = (logits, layout.n_pad, offset, spec.out_features)
            # dependency probabilities for later levels: one launch per source buffer
            by_buffer: Dict[int, List[str]] = {}
            for name in layout.feeds_later:
                by_buffer.setdefault(id(produced[name][0]), []).append(name)
            for names in by_buffer.values():
                buffer, ld = produced[names[0]][0], produced[names[0]][1]
                columns, widths_list, targets = [], [], []
                for name in names:
                    _, _, column, width = produced[name]
                    target_column, expected = self.dep_cols[name]
                    if width - skip != expected:
                        raise ValueError(
                            f"classifier {name!r} produces {width - skip} dependency features but its dependents expect {expected}"
                        )
                    columns.append(column)
                    widths_list.append(width)
                    targets.append(target_column)
                col_off = self._int_tensor(("dep_col", tuple(columns)), columns, device, torch.int32)
                widths = self._int_tensor(("dep_w", tuple(widths_list)), widths_list, device, torch.int32)
                dst_col = self._int_tensor(("dep_dst", tuple(targets)), targets, device, torch.int32)
                ops.dependency_softmax(buffer, ld, rows, col_off, widths, dst_col, len(names), skip, plan.x, self.ldx)
            produced_all.update(produced)

        # outputs in the reference's (topological) order, independent of the column layout
        for spec in projection._specs:
            classifier = projection._layers[spec.name]
            buffer, ld, column, width = produced_all[spec.name]
            if classifier._allophone_layer is not None:
                if predict:
                    # acoustic_model.py:164-166: phone logits under both names
                    heads.append((ProjectionEntryConfig.PHONE, buffer, ld, column, width))
                    heads.append((ProjectionEntryConfig.PHONEME_LAYER, buffer, ld, column, width))
                else:
                    mapped, argmax = self._map_allophones(buffer[:, column : column + width].view(n_utt, seq, width), batch.language_ids)
                    if keep is not None:
                        keep["allophone"] = dict(argmax=argmax, source=(buffer, ld, column, width), name=spec.name)
                    heads.append((ProjectionEntryConfig.PHONEME_LAYER, mapped.view(rows, -1), mapped.shape[-1], 0, mapped.shape[-1]))
            else:
                heads.append((spec.name, buffer, ld, column, width))

        if keep is not None:
            keep.update(level_logits=level_logits, level_bf16=level_bf16, produced=produced_all, composed=composed_info)
        return heads, produced_all

    # ------------------------------------------------------------------ forward
    def forward(self, batch: Batch, target_feature_indices: Optional[Tensor], predict: bool, log_probabilities: bool):
        from .network.acoustic_model import Predictions

        model = self.model
        projection = model._projection
        if not self._layout_ready:
            self._build_layout()
        acoustic = model._acoustic_model
        if torch.is_grad_enabled() and not log_probabilities and not hasattr(acoustic, "_model"):
            # from-scratch transformer encoder (network/transformer.py): its own plan type with its own backward pass
            need_encoder = any(p.requires_grad for p in acoustic.parameters())
            if need_encoder or any(p.requires_grad for p in projection.parameters()):
                return self._forward_differentiable(batch, target_feature_indices, predict, need_encoder, False)
        elif torch.is_grad_enabled() and not log_probabilities:
            weights = acoustic._model
            need_extractor = any(p.requires_grad for p in weights.feature_extractor.parameters())
            need_encoder = any(p.requires_grad for p in weights.encoder.parameters())
            need_feature_projection = any(p.requires_grad for p in weights.feature_projection.parameters())
            if need_encoder or need_feature_projection or need_extractor or any(p.requires_grad for p in projection.parameters()):
                return self._forward_differentiable(batch, target_feature_indices, predict, need_encoder, need_feature_projection, need_extractor)
        plan, frames = acoustic.encode(batch, self.ldx, self.hidden_blocks)
        device = plan.x.device
        self._ensure_weights(device)
        rows, n_utt, seq = plan.rows, plan.n_utt, plan.seq
        heads, _ = self._run_levels(plan, batch, target_feature_indices, predict, keep=None)

        # frames the caller sees: fewer than the launch list's when the batch was padded up to a length bucket
        shown = getattr(plan, "seq_out", None) or seq
        if not log_probabilities:
            outputs = {
                name: buffer[:, column : column + width].reshape(n_utt, seq, width).transpose(0, 1)[:shown]
                for name, buffer, ld, column, width in heads
            }
            return Predictions(outputs, frames)

        # fused per-head log_softmax (+ argmax / max log-prob for greedy decoding)
        total = sum(width for _, _, _, _, width in heads) * rows
        out = torch.empty(total, device=device, dtype=torch.float32)
        n_heads = len(heads)
        argmax = torch.empty(n_heads, rows, device=device, dtype=torch.int32)
        maxlp = torch.empty(n_heads, rows, device=device, dtype=torch.float32)
        outputs: Dict[str, Tensor] = {}
        out_offsets: List[int] = []
        position = 0
        for name, buffer, ld, column, width in heads:
            out_offsets.append(position)
            outputs[name] = out[position : position + rows * width].view(n_utt, seq, width).transpose(0, 1)[:shown]
            position += rows * width
        # group heads by source buffer; narrow heads of one buffer go into a single launch
        index = 0
        while index < n_heads:
            buffer, ld = heads[index][1], heads[index][2]
            end = index
            while end < n_heads and heads[end][1] is buffer:
                end += 1
            group = list(range(index, end))
            narrow = [h for h in group if heads[h][4] <= 128]
            wide = [h for h in group if heads[h][4] > 128]
            if narrow:
                first, last = narrow[0], narrow[-1]
                lo = min(heads[h][3] for h in narrow) // 4 * 4
                hi = _round_up(max(heads[h][3] + heads[h][4] for h in narrow), 4)
                hi = min(hi, ld)
                contiguous = narrow == list(range(first, last + 1))
                if (hi - lo) * 4 * 32 > 190 * 1024 or not contiguous:
                    wide = sorted(wide + narrow)
                    narrow = []
                else:
                    cols = [heads[h][3] for h in narrow]
                    wids = [heads[h][4] for h in narrow]
                    outs = [out_offsets[h] for h in narrow]
                    col_off = self._int_tensor(("lsm_col", tuple(cols)), cols, device, torch.int32)
                    widths = self._int_tensor(("lsm_w", tuple(wids)), wids, device, torch.int32)
                    offs = self._int_tensor(("lsm_out", tuple(outs)), outs, device, torch.int64)
                    ops.log_softmax_heads(
                        buffer, ld, rows, lo, hi - lo, col_off, widths, offs, len(narrow), out, argmax[first : last + 1], maxlp[first : last + 1]
                    )
            for h in wide:
                _, _, _, column, width = heads[h]
                ops.log_softmax_wide(
                    buffer[:, column:], ld, rows, width, out[out_offsets[h] :], width, argmax[h], maxlp[h]
                )
            index = end
        predictions = Predictions(outputs, frames)
        predictions._decode_cache = dict(  # type: ignore[attr-defined]
            argmax=argmax,
            maxlp=maxlp,
            head_index={name: h for h, (name, *_rest) in enumerate(heads)},
            frames32=plan.frames32.clone(),
            n_utt=n_utt,
            seq=seq,
            flat=out,  # the one buffer every head's log-probabilities are views of (Estimator._predict_graphed copies it)
        )
        return predictions

    def _encoder_prefix(self) -> str:
        """``state_dict`` prefix of the encoder parameters the plan's ``backward`` names its gradients by."""
        return "_acoustic_model._model." if hasattr(self.model._acoustic_model, "_model") else "_acoustic_model."

    @staticmethod
    def _rank() -> int:
        import torch.distributed as dist

        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    # ------------------------------------------------------------------ differentiable path (training)
    def _forward_differentiable(
        self, batch: Batch, target_feature_indices: Optional[Tensor], predict: bool, need_encoder: bool, need_feature_projection: bool,
        need_extractor: bool = False,
    ):  # fmt: skip
        """Training / validation forward with autograd: the whole model is ONE ``torch.autograd.Function``.

        torch only carries the graph edge from the returned logits back to the parameters; forward and
        backward run in ``liballophant_b200.so`` (``EncoderPlan.run`` / ``EncoderPlan.backward`` for the
        wav2vec2 encoder, the level GEMMs and the small kernels of ``aph_train.cu`` for the heads).
        In ``train()`` mode the stochastic regularisation of the reference is applied — HF dropout / LayerDrop /
        SpecAugment inside the encoder (``engine.Stochastic``) and the dropout on the acoustic-model outputs
        (``acoustic_model.py:486-488``) — with counter-based masks seeded from torch's CPU generator; in ``eval()``
        mode the arithmetic is deterministic, which is what the parity tests pin against the reference."""
        from .engine import Stochastic
        from .network.acoustic_model import Predictions

        acoustic = self.model._acoustic_model
        through_encoder = need_encoder or need_feature_projection or need_extractor
        stochastic = None
        input_dropout: Dict[int, ops.Dropout] = {}  # column of X -> dropout of that classifier input block
        if self.model.training:
            seed = int(torch.randint(0, 2**31 - 1, (1,))) ^ (self._rank() * 0x9E3779B1 & 0x7FFFFFFF)
            if hasattr(acoustic, "_model") and acoustic._model.training:  # HF regularises by module mode, also when the encoder is frozen
                stochastic = Stochastic.from_config(acoustic._model.config, seed)
                stochastic.skip_layers = self.skip_layers_override
            elif not hasattr(acoustic, "_model") and acoustic.training:
                stochastic = seed  # the from-scratch transformer encoder takes the seed of its dropout masks
            rate = self.model._projection._acoustic_model_dropout
            if rate is not None and rate.p > 0:
                blocks = {0: -1, **{column: index for index, column in self.hidden_blocks.items()}}
                if sum(1 for name, column in self.x_cols.items() if column == 0) > 1:
                    raise NotImplementedError("acoustic_model_dropout with both OUTPUT and the last OUTPUT_i as dependencies")
                input_dropout = {column: ops.Dropout.site(rate.p, seed, Stochastic.SITE_CLASSIFIER_INPUT + index + 1) for column, index in blocks.items()}
        with torch.no_grad():
            extra = dict(train_extractor=True) if need_extractor else {}
            plan, frames = acoustic.encode(batch, self.ldx, self.hidden_blocks, training=through_encoder or stochastic is not None, stochastic=stochastic, **extra)
            self._ensure_weights(plan.x.device)
            hidden = self.model._projection._output_features
            for column, drop in input_dropout.items():
                ops.dropout_bf16_2d(plan.x[:, column:], self.ldx, plan.rows, hidden, drop)
        self.last_regularisation = dict(stochastic=stochastic, input_dropout=input_dropout, plan=plan)  # introspection (tests)
        cached = getattr(self, "_named_parameters", None)
        if cached is None:  # nn.Module.named_parameters walks ~500 tensors through several generators: 1.7 ms per step
            weights = acoustic._model if hasattr(acoustic, "_model") else acoustic
            cached = self._named_parameters = (
                [(f"_projection.{name}", parameter) for name, parameter in self.model._projection.named_parameters()],
                [(self._encoder_prefix() + name, parameter) for name, parameter in weights.named_parameters()],
            )
        named = list(cached[0])
        if through_encoder:
            named += [item for item in cached[1] if item[1].requires_grad]
        state: Dict[str, Any] = dict(
            batch=batch, tfi=target_feature_indices, predict=predict, plan=plan, names=[n for n, _ in named], generation=None,
            need_encoder=need_encoder, need_feature_projection=need_feature_projection, need_extractor=need_extractor, input_dropout=input_dropout,
        )  # fmt: skip
        outputs = _ProjectionFunction.apply(self, state, *[p for _, p in named])
        names = state["head_names"]
        return Predictions({name: tensor.transpose(0, 1) for name, tensor in zip(names, outputs)}, frames)

    @torch.no_grad()
    def _differentiable_forward_impl(self, state: Dict[str, Any]) -> Tuple[Tensor, ...]:
        plan = state["plan"]
        keep: Dict[str, Any] = {}
        heads, _ = self._run_levels(plan, state["batch"], state["tfi"], state["predict"], keep)
        rows, n_utt, seq = plan.rows, plan.n_utt, plan.seq
        state.update(keep)
        state["generation"] = plan.generation  # the plan's buffers (X, kept activations) are overwritten by its next run
        state["heads"] = heads
        state["head_names"] = [name for name, *_ in heads]
        state["dims"] = (rows, n_utt, seq)
        # every classifier's logits as a tensor of its own (autograd hands each to the caller's loss): one flat buffer, one launch
        flat = torch.empty(sum(width for *_, width in heads) * rows, device=plan.x.device, dtype=torch.float32)
        outputs, src, dst, position = [], [], [], 0
        for _, buffer, ld, column, width in heads:
            block = flat[position : position + rows * width]
            position += rows * width
            src.append((buffer, ld, column, width))
            dst.append((block, width, 0, width))
            outputs.append(block.view(n_utt, seq, width))
        ops.copy_head_blocks(src, dst, rows)
        return tuple(outputs)

    @torch.no_grad()
    def _differentiable_backward_impl(self, state: Dict[str, Any], grads: Tuple[Optional[Tensor], ...]) -> List[Optional[Tensor]]:
        projection = self.model._projection
        plan = state["plan"]
        if plan.generation != state["generation"]:
            raise RuntimeError(
                "allophant_b200: the activations of this forward pass were overwritten by a later forward pass of the same "
                "batch shape; call backward() before running the next training forward"
            )
        rows, n_utt, seq = state["dims"]
        x = plan.x
        device = x.device
        skip = 0 if projection._dependency_blanks else projection._blank_offset
        param_grads: Dict[str, Tensor] = {}
        through_encoder = state["need_encoder"] or state["need_feature_projection"] or state.get("need_extractor", False)
        need_dx = through_encoder or any(layout.feeds_later for layout in self.levels)

        # gradient of every classifier's final logits, keyed by classifier name, fp32 [rows, width]
        head_grads: Dict[str, Tensor] = {}
        for (name, _, _, _, width), grad in zip(state["heads"], grads):
            if grad is None:
                continue
            flat = grad.float().reshape(rows, width)
            head_grads[name] = head_grads[name] + flat if name in head_grads else flat
        # allophone layer: "phoneme" (mapped) -> gradient w.r.t. the phone logits and the matrices
        allophone = state.get("allophone")
        if allophone is not None and ProjectionEntryConfig.PHONEME_LAYER in head_grads:
            buffer, ld, column, width = allophone["source"]
            layer = projection._layers[allophone["name"]]._allophone_layer
            languages = state["batch"].language_ids.to(device=device, dtype=torch.int64).contiguous()
            grad_mapped = head_grads.pop(ProjectionEntryConfig.PHONEME_LAYER).contiguous()
            grad_phone, grad_matrices = ops.allophone_backward(
                grad_mapped.view(n_utt, seq, -1), allophone["argmax"], buffer[:, column : column + width].view(n_utt, seq, width),
                layer._allophone_matrices.detach().float().contiguous(), languages, True, True,
            )  # fmt: skip
            param_grads[f"_projection._layers.{allophone['name']}._allophone_layer._allophone_matrices"] = grad_matrices
            head_grads["__phone__" + allophone["name"]] = grad_phone.view(rows, width)
        elif state["predict"]:
            # predict=True returns the phone logits under both names: both gradients flow into the same logits
            phone = head_grads.pop(ProjectionEntryConfig.PHONE, None)
            if phone is not None:
                key = ProjectionEntryConfig.PHONEME_LAYER
                head_grads[key] = head_grads[key] + phone if key in head_grads else phone

        d_x: Optional[Tensor] = None  # dL/dX fp32 [rows, ldx], accumulated over the levels (highest level first)
        for level_index in reversed(range(len(self.levels))):
            layout = self.levels[level_index]
            n_pad = layout.n_pad
            grad_level = torch.zeros(rows, n_pad, device=device, dtype=torch.float32)
            # gradients arriving through the dependency probabilities of later levels (all processed already)
            feeders = [name for name in layout.feeds_later]
            if feeders and d_x is not None:
                for name in feeders:
                    if projection._layers[name]._composition_layer is not None or projection._layers[name]._allophone_layer is not None:
                        raise NotImplementedError("gradients through a composed/allophone phoneme layer used as a dependency are not implemented")
                x_col = self._int_tensor(("bwd_xcol", level_index), [self.dep_cols[name][0] for name in feeders], device, torch.int32)
                widths = self._int_tensor(("bwd_w", level_index, skip), [self.dep_cols[name][1] + skip for name in feeders], device, torch.int32)
                dst_col = self._int_tensor(("bwd_dst", level_index), [layout.offsets[name] for name in feeders], device, torch.int32)
                ops.softmax_backward_cols(d_x, self.ldx, x, self.ldx, rows, x_col, widths, dst_col, len(feeders), skip, grad_level, n_pad)
            # plain linear heads: their logits gradients join the level's gradient matrix in ONE launch (not one strided add each)
            plain_src, plain_dst = [], []
            for spec in layout.specs:
                if projection._layers[spec.name]._composition_layer is not None:
                    continue
                grad = head_grads.get(spec.name)
                if grad is None:
                    grad = head_grads.get("__phone__" + spec.name)
                if grad is None:
                    continue
                grad = grad if grad.is_contiguous() else grad.contiguous()
                plain_src.append((grad, spec.out_features, 0, spec.out_features))
                plain_dst.append((grad_level, n_pad, layout.offsets[spec.name], spec.out_features))
            ops.copy_head_blocks(plain_src, plain_dst, rows, accumulate=True)
            for spec in layout.specs:
                classifier = projection._layers[spec.name]
                offset = layout.offsets[spec.name]
                grad = head_grads.get(spec.name)
                if grad is None:
                    grad = head_grads.get("__phone__" + spec.name)
                if grad is None:
                    continue
                if classifier._composition_layer is None:
                    continue
                # composed phoneme logits = (projection @ table^T) / sqrt(E)   (acoustic_model.py:219-234)
                info = state["composed"][spec.name]
                table, classes = info["table"], info["classes"]
                embedding = classifier._composition_layer.embedding_size
                v_pad = table.shape[0]
                grad_logits = torch.zeros(rows, v_pad, device=device, dtype=torch.bfloat16)
                grad_logits[:, :classes] = grad
                scale = 1.0 / math.sqrt(embedding)
                # d projection [rows, E] = d logits @ table * scale, accumulated into this level's logits gradient
                grad_projection = torch.empty(rows, embedding, device=device, dtype=torch.float32)
                ops.run_gemm(ops.make_dgrad_args(grad_logits, table, rows=rows, ld_dy=v_pad, k=v_pad, n=embedding, ld_w=embedding, scale=scale,
                                                 out_f32=grad_projection, ld_f32=embedding))  # fmt: skip
                grad_level[:, offset : offset + embedding] += grad_projection
                # d table [v_pad, E] = d logits^T @ projection * scale -> EmbeddingBag backward
                projection_bf16 = state["level_bf16"][level_index]
                grad_table = torch.empty(v_pad, embedding, device=device, dtype=torch.float32)
                ops.run_gemm(ops.make_wgrad_args(grad_logits, projection_bf16[:, offset:], grad_table, rows=rows, m=v_pad, ld_dy=v_pad,
                                                 n=embedding, ld_x=n_pad, ld_out=embedding, scale=scale))  # fmt: skip
                layer = classifier._composition_layer
                weight = layer._attribute_embeddings.weight
                grad_weight = torch.zeros(weight.shape, device=device, dtype=torch.float32)
                if state["tfi"] is None:
                    indices, offsets = layer._dense_feature_table, None
                else:
                    indices, offsets = state["tfi"], layer._category_offsets
                indices = indices.to(device=device, dtype=torch.int64).contiguous()
                offsets = None if offsets is None else offsets.to(device=device, dtype=torch.int64).contiguous().view(-1)
                ops.embedding_bag_backward(grad_table, embedding, indices, offsets, grad_weight)
                param_grads[f"_projection._layers.{spec.name}._composition_layer._attribute_embeddings.weight"] = grad_weight
            # weight / bias gradients of the level: dW = dY^T X, db = colsum(dY)
            grad_level_bf16 = ops.cast_bf16(grad_level)
            grad_weight_level = torch.empty(n_pad, self.ldx, device=device, dtype=torch.float32)
            ops.run_gemm(ops.make_wgrad_args(grad_level_bf16, x, grad_weight_level, rows=rows, m=n_pad, ld_dy=n_pad, n=self.ldx, ld_x=self.ldx,
                                             ld_out=self.ldx))  # fmt: skip
            grad_bias_level = ops.colsum_f32(grad_level, rows, n_pad, n_pad)
            for spec in layout.specs:
                linear = projection._layers[spec.name]._time_distributed_layer
                row = layout.offsets[spec.name]
                grad_w = torch.empty(linear.weight.shape, device=device, dtype=torch.float32)
                source_column = 0
                for dependency in spec.dependencies:
                    if _PATTERN.match(dependency.name):
                        target = self.x_cols[dependency.name]
                    else:
                        target = self.dep_cols[dependency.name][0]
                    grad_w[:, source_column : source_column + dependency.size] = grad_weight_level[row : row + spec.out_features, target : target + dependency.size]
                    source_column += dependency.size
                param_grads[f"_projection._layers.{spec.name}._time_distributed_layer.weight"] = grad_w
                param_grads[f"_projection._layers.{spec.name}._time_distributed_layer.bias"] = grad_bias_level[row : row + spec.out_features].clone()
            # data gradient dX += dY @ W_level (W read in its forward layout as an MN-major operand)
            if need_dx and (through_encoder or level_index > 0):
                if d_x is None:
                    d_x = torch.empty(rows, self.ldx, device=device, dtype=torch.float32)
                    resid = None
                else:
                    resid = d_x
                ops.run_gemm(ops.make_dgrad_args(grad_level_bf16, self.level_w[level_index], rows=rows, ld_dy=n_pad, k=n_pad, n=self.ldx, ld_w=self.ldx,
                                                 resid=resid, ld_resid=self.ldx, out_f32=d_x, ld_f32=self.ldx))  # fmt: skip

        reducer = self.gradient_reducer
        if reducer is not None:
            param_grads = reducer.submit_tensors(param_grads)  # the heads' gradients travel while the encoder backward runs
        if through_encoder:
            assert d_x is not None
            hidden = projection._output_features
            for column, drop in state["input_dropout"].items():  # acoustic_model.py:486-488, same masks as the forward
                block = d_x[:, column:]
                ops.dropout_2d(block, d_x.shape[1], rows, hidden, drop, out_f32=block, ld_f32=d_x.shape[1])
            extra = dict(need_extractor=True) if state.get("need_extractor") else {}
            encoder_grads = plan.backward(
                d_x, state["need_encoder"], state["need_feature_projection"], None if reducer is None else reducer.submit, **extra
            )
            for name, value in encoder_grads.items():
                param_grads[self._encoder_prefix() + name] = value
        if reducer is not None:
            reducer.finish()
        return [param_grads.get(name) for name in state["names"]]

    # ------------------------------------------------------------------ allophone layer
    def _allophone_csr(self, layer: Any, device: torch.device) -> Tuple[Tensor, Tensor]:
        """CSR list of the allowed (phone, phoneme) pairs per language; the mask is fixed at construction
        (``acoustic_model.py:134-136``), only the matrix values train."""
        cached = getattr(self, "_csr_cache", None)
        if cached is not None and cached[0] == (layer._allophone_mask.data_ptr(), str(device)):
            return cached[1], cached[2]
        allowed = ~layer._allophone_mask.cpu()  # [L, P+1, Q+1]
        offsets, phones = [0], []
        for language in range(allowed.shape[0]):
            for phoneme in range(allowed.shape[2]):
                listed = torch.nonzero(allowed[language, :, phoneme]).flatten().tolist()
                phones.extend(listed)
                offsets.append(len(phones))
        csr_off = torch.tensor(offsets, dtype=torch.int32, device=device)
        csr_p = torch.tensor(phones if phones else [0], dtype=torch.int32, device=device)
        self._csr_cache = ((layer._allophone_mask.data_ptr(), str(device)), csr_off, csr_p)
        return csr_off, csr_p

    def _map_allophones(self, phone_logits: Tensor, language_ids: Tensor) -> Tuple[Tensor, Tensor]:
        """``phone_logits`` fp32 [N, T', P+1] (batch-first view) -> ([N, T', Q+1], argmax)."""
        layer = self.model._projection._layers[ProjectionEntryConfig.PHONEME_LAYER]._allophone_layer
        if layer is None:
            raise ValueError("Can't map phones to allophones with a model without an allophone layer")
        device = phone_logits.device
        csr_off, csr_p = self._allophone_csr(layer, device)
        languages = language_ids.to(device=device, dtype=torch.int64).contiguous()
        matrices = layer._allophone_matrices.detach().float().contiguous()
        return ops.allophone_forward(phone_logits, matrices, csr_off, csr_p, languages)

    def map_allophones(self, phone_logits: Tensor, language_ids: Tensor) -> Tensor:
        """``Allophant.map_allophones`` (``acoustic_model.py:1040-1041``): time-first ``[T', N, P+1]`` in and out."""
        if not phone_logits.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only")
        batch_first = phone_logits.float().transpose(0, 1)
        if batch_first.stride(-1) != 1:
            batch_first = batch_first.contiguous()
        mapped, _ = self._map_allophones(batch_first, language_ids)
        return mapped.transpose(0, 1)
