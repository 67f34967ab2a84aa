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
from typing import Any, Dict, List, Optional, Tuple

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


class HeadsRuntime:
    def __init__(self, model: Any) -> None:
        self.model = model
        self._layout_ready = False
        self._weights_version: Optional[Tuple[int, ...]] = None
        self._composed_cache: Dict[Tuple[Any, ...], Tuple[Tensor, int]] = {}
        self._index_cache: Dict[Tuple[Any, ...], Dict[str, Tensor]] = {}

    # ------------------------------------------------------------------ static layout
    def _build_layout(self) -> None:
        projection = self.model._projection
        hidden = projection._output_features
        n_layers = self.model._acoustic_model._model.config.num_hidden_layers
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
        params = list(self.model._projection.parameters())
        return tuple(_version(p) for p in params) + tuple(p.data_ptr() for p in params)

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
            for spec in layout.specs:
                linear = projection._layers[spec.name]._time_distributed_layer
                if not linear.weight.is_cuda:
                    raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
                row = layout.offsets[spec.name]
                source_column = 0
                for dependency in spec.dependencies:
                    if _PATTERN.match(dependency.name):
                        target = self.x_cols[dependency.name]
                    else:
                        target = self.dep_cols[dependency.name][0]
                    weight[row : row + spec.out_features, target : target + dependency.size] += linear.weight.detach()[
                        :, source_column : source_column + dependency.size
                    ].float()
                    source_column += dependency.size
                bias[row : row + spec.out_features] = linear.bias.detach().float()
            self.level_w.append(ops.cast_bf16(weight))
            self.level_b.append(bias)
        self._composed_cache.clear()
        self._weights_version = version

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
        if int(err.item()) != 0:  # one-time check per inventory (EmbeddingBag raises on out-of-range indices too)
            raise IndexError("target_feature_indices contain a category outside the attribute embedding table")
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

    # ------------------------------------------------------------------ forward
    def forward(self, batch: Batch, target_feature_indices: Optional[Tensor], predict: bool, log_probabilities: bool):
        from .network.acoustic_model import Predictions

        model = self.model
        projection = model._projection
        if not self._layout_ready:
            self._build_layout()
        if torch.is_grad_enabled() and model.training and any(p.requires_grad for p in model.parameters()):
            raise NotImplementedError(
                "allophant_b200: the differentiable training forward is not available in this build; "
                "wrap inference in torch.inference_mode()/no_grad() or call model.eval()"
            )
        acoustic = model._acoustic_model
        plan, frames = acoustic.encode(batch, self.ldx, self.hidden_blocks)
        device = plan.x.device
        self._ensure_weights(device)
        rows, n_utt, seq = plan.rows, plan.n_utt, plan.seq
        skip = 0 if projection._dependency_blanks else projection._blank_offset

        # (name, logits buffer, leading dim, first column, width) in topological order
        heads: List[Tuple[str, Tensor, int, int, int]] = []
        level_logits: List[Tensor] = []
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
            produced: Dict[str, Tuple[Tensor, int, int, int]] = {}
            for spec in layout.specs:
                classifier = projection._layers[spec.name]
                offset = layout.offsets[spec.name]
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
                    level_logits.append(composed)
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
                if not predict:
                    raise NotImplementedError(
                        "allophant_b200: the allophone layer's training forward (map_allophones) is not available in this build"
                    )
                heads.append((ProjectionEntryConfig.PHONE, buffer, ld, column, width))
                heads.append((ProjectionEntryConfig.PHONEME_LAYER, buffer, ld, column, width))
            else:
                heads.append((spec.name, buffer, ld, column, width))

        if not log_probabilities:
            outputs = {
                name: buffer[:, column : column + width].reshape(n_utt, seq, width).transpose(0, 1)
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
            outputs[name] = out[position : position + rows * width].view(n_utt, seq, width).transpose(0, 1)
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
        )
        return predictions

    def map_allophones(self, phone_logits: Tensor, language_ids: Tensor) -> Tensor:
        layer = self.model._projection._layers[ProjectionEntryConfig.PHONEME_LAYER]._allophone_layer
        if layer is None:
            raise ValueError("Can't map phones to allophones with a model without an allophone layer")
        raise NotImplementedError("allophant_b200: map_allophones is not available in this build")
