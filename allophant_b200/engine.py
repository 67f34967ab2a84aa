"""Host-side orchestration of the CUDA encoder: packed weights, workspaces, launch sequence.

One ``EncoderPlan`` exists per (batch size, padded sample count) and owns every intermediate
buffer in HBM, laid out channels-last / row-major so each kernel streams contiguous rows:

    audio fp32 [N,T] ──conv0+LN+GELU──▶ bf16 [N,L0,512] ──6× (implicit-GEMM conv, LN+GELU)──▶ bf16 [N,T',512]
      ──LN──▶ bf16 [M,512] ──GEMM(+mask)──▶ hidden fp32 [M,1024] (+ bf16 copy for the pos-conv taps)
      ──pos-conv GEMM(+GELU+residual)──▶ hidden ──24× {LN, QKV GEMM, attention, out-proj(+res), LN,
      FFN1(+GELU), FFN2(+res)}──▶ final LN ──▶ bf16 feature matrix X [M, ldx]          (M = N·T')

The residual stream stays fp32; every GEMM operand is bf16 with fp32 accumulation in TMEM.
No kernel reads a dense attention mask: frame counts per utterance are computed on the device
and passed to the masking epilogue and the attention kernel.  Nothing in here synchronises
with the host.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib, ops
from .network.wav2vec2 import Wav2Vec2EncoderConfig, Wav2Vec2Weights


# Bumped by optimisers that update parameters through raw pointers (allophant_b200.optim.FusedAdam): torch's version
# counters do not see such writes, so the packed bf16 operands key on this number as well.
_WEIGHT_GENERATION = 0


_BACKWARD_OVERLAP = True


def set_backward_overlap(enabled: bool) -> bool:
    """Process-wide switch of the second stream of ``EncoderPlan.backward`` (weight / bias gradients off the data-gradient
    path; on by default, ``APH_BWD_OVERLAP=0`` disables it when a plan is built).  Returns the previous setting.  Measurement
    code turns it off to time kernels one at a time: records of overlapping kernels include the time they share the SMs."""
    global _BACKWARD_OVERLAP
    before, _BACKWARD_OVERLAP = _BACKWARD_OVERLAP, bool(enabled)
    return before


def bump_weight_generation() -> None:
    global _WEIGHT_GENERATION
    _WEIGHT_GENERATION += 1


def weight_generation() -> int:
    return _WEIGHT_GENERATION


def conv_out_length(length: int, kernel: int, stride: int) -> int:
    return (length - kernel) // stride + 1


def frames_for(samples: int, cfg: Wav2Vec2EncoderConfig) -> List[int]:
    """Output length of every conv layer for ``samples`` input samples."""
    lengths = []
    for kernel, stride in zip(cfg.conv_kernel, cfg.conv_stride):
        samples = conv_out_length(samples, kernel, stride)
        lengths.append(samples)
    return lengths


def min_samples_for(frames: int, cfg: Wav2Vec2EncoderConfig) -> int:
    """Smallest sample count whose convolutional feature extractor output has ``frames`` frames."""
    length = frames
    for kernel, stride in zip(reversed(cfg.conv_kernel), reversed(cfg.conv_stride)):
        length = (length - 1) * stride + kernel
    return length


def bucket_samples(samples: int, cfg: Wav2Vec2EncoderConfig, frame_multiple: int) -> Tuple[int, int]:
    """``(padded sample count, frames of the unpadded input)``: inputs are padded (with silence, behind every utterance's own
    length) up to the LARGEST sample count whose frame count is the next multiple of ``frame_multiple``, so that a stream of
    ragged batches (``MaxFrameBatchSampler``, ``batching.py:94-139``) maps onto a few launch lists instead of one per length."""
    frames = frames_for(samples, cfg)[-1]
    if frame_multiple <= 1 or frames < 1:
        return samples, frames
    target = -(-frames // frame_multiple) * frame_multiple
    return max(samples, min_samples_for(target + 1, cfg) - 1), frames


class WorkspaceArena:
    """One grow-only device buffer that the inference launch lists of a model overlay (each ``EncoderPlan`` carves its
    workspaces from offset 0): a ragged stream switches between launch lists without allocating.  Everything a plan reads
    before writing is either rewritten or cleared at the start of each run (``EncoderPlan.run``)."""

    ALIGN = 1024

    def __init__(self) -> None:
        self.buffer: Optional[Tensor] = None
        self.generation = 0  # bumped when the buffer is replaced: every plan carved from the old one is void

    def capacity(self) -> int:
        return 0 if self.buffer is None else self.buffer.numel()

    def reserve(self, n_bytes: int, device: torch.device) -> None:
        if self.capacity() < n_bytes:
            self.buffer = None  # release first: two generations of a multi-GB arena need not coexist
            self.buffer = torch.empty(n_bytes, device=device, dtype=torch.uint8)
            self.generation += 1

    def carver(self) -> "ArenaCarver":
        return ArenaCarver(self)


class ArenaCarver:
    """Bump allocation over a ``WorkspaceArena``; requests beyond its capacity are served by ordinary allocations and counted,
    so the caller learns the size to reserve and rebuilds."""

    def __init__(self, arena: WorkspaceArena) -> None:
        self.arena = arena
        self.offset = 0
        self.overflow = False

    def zeros(self, *shape: int, device: torch.device, dtype: torch.dtype) -> Tensor:
        numel = 1
        for extent in shape:
            numel *= int(extent)
        n_bytes = numel * torch.empty((), dtype=dtype).element_size()
        start = self.offset
        self.offset = (start + n_bytes + WorkspaceArena.ALIGN - 1) // WorkspaceArena.ALIGN * WorkspaceArena.ALIGN
        if self.arena.buffer is None or self.offset > self.arena.capacity():
            self.overflow = True
            return torch.zeros(*shape, device=device, dtype=dtype)
        view = self.arena.buffer[start : start + n_bytes].view(dtype).view(*shape)
        view.zero_()
        return view


class PackedEncoder:
    """bf16 GEMM operands packed from the fp32 master parameters.

    The operand buffers are allocated once per parameter layout (``layout_id``) and REFILLED in place when the
    parameters change, so the launch lists of the ``EncoderPlan``s (which hold raw pointers) stay valid across
    optimiser steps.  With ``allophant_b200.optim.FusedAdam`` the Linear operands are not even refilled: the Adam
    kernel writes the bf16 copy itself (``shadow_map``), and ``after_fused_step`` refreshes the few derived tensors
    (fused QKV bias, weight-normed positional conv)."""

    def __init__(self, weights: Wav2Vec2Weights) -> None:
        self.weights = weights
        self.cfg = weights.config
        self._version: Optional[Tuple[int, ...]] = None
        self._params: Optional[List[Tensor]] = None
        self._pointers: Optional[Tuple[int, ...]] = None
        self.device: Optional[torch.device] = None
        self.layout_id = 0
        self._folded: Optional[List[Dict[str, Tensor]]] = None
        self._folded_layout = -1
        self._folded_version: Optional[Tuple[int, ...]] = None

    def _current_version(self) -> Tuple[int, ...]:
        def version(p: Tensor) -> int:
            try:
                return p._version
            except RuntimeError:  # inference tensors do not track a version counter
                return -1

        params = self._params
        if params is None:  # the module tree is fixed after construction: walk it once
            params = self._params = list(self.weights.parameters())
        return (weight_generation(),) + tuple(version(p) for p in params) + tuple(p.data_ptr() for p in params)

    def ensure(self) -> None:
        version = self._current_version()
        if version != self._version:
            pointers = tuple(p.data_ptr() for p in self._params or [])
            if pointers != self._pointers:  # first use, or the parameters moved (``.to(device)``, ``load_state_dict`` keeps them)
                self._allocate()
                self._pointers = pointers
            self._fill()
            self._version = version

    @torch.no_grad()
    def _allocate(self) -> None:
        w, cfg = self.weights, self.cfg
        first = next(w.parameters())
        if not first.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
        self.device = dev = first.device
        if cfg.hidden_size % 256 != 0 or cfg.head_dim != 64:
            raise NotImplementedError("the CUDA encoder supports head_dim 64 and hidden sizes that are multiples of 256")
        if any(c != 512 for c in cfg.conv_dim) or cfg.conv_kernel[0] != 10 or cfg.conv_stride[0] != 5:
            raise NotImplementedError("the CUDA feature extractor supports the wav2vec2 layout (7 x 512 channels, k0=10, s0=5)")

        def f32(p: Tensor) -> Tensor:  # fp32 contiguous parameters are used in place (no copy): always current
            return p.detach().float().contiguous()

        bf16 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)  # noqa: E731
        H, FF = cfg.hidden_size, cfg.intermediate_size
        fe = w.feature_extractor.conv_layers
        self.conv0_w = f32(fe[0].conv.weight).view(512, 10)
        self.conv_bias = [f32(l.conv.bias) if l.conv.bias is not None else None for l in fe]
        self.conv_ln = [(f32(l.layer_norm.weight), f32(l.layer_norm.bias)) if hasattr(l, "layer_norm") else None for l in fe]
        self.conv_w = [None] + [bf16(l.conv.weight.shape[0], l.conv.weight.shape[2] * l.conv.weight.shape[1]) for l in fe[1:]]
        fp = w.feature_projection
        self.fp_ln = (f32(fp.layer_norm.weight), f32(fp.layer_norm.bias))
        self.fp_w = bf16(H, 512)
        self.fp_b = f32(fp.projection.bias)
        pc = w.encoder.pos_conv_embed.conv
        v = pc.parametrizations.weight.original1
        # Groups that are not 64 channels wide (wav2vec2-base: 16 groups of 48): the tap GEMM contracts every 64-column output
        # tile over a block-diagonal super group of `pos_span` channels (the smallest common multiple of the group width and 64)
        # and the operand carries zeros where input and output channel belong to different groups (aph_gemm_args.taps_span).
        group = v.shape[1]
        self.pos_span = 64 if group == 64 else group * 64 // math.gcd(group, 64)
        if H % self.pos_span != 0:
            raise NotImplementedError(f"positional conv groups of {group} channels do not tile the hidden size {H}")
        self.pos_w = bf16(v.shape[0], v.shape[2] * self.pos_span)
        self.pos_w_grouped = self.pos_w if group == 64 else bf16(v.shape[0], v.shape[2] * group)
        self.pos_b = f32(pc.bias)
        self.layers = []
        for layer in w.encoder.layers:
            att = layer.attention
            self.layers.append(
                dict(
                    ln1=(f32(layer.layer_norm.weight), f32(layer.layer_norm.bias)),
                    wqkv=bf16(3 * H, H),
                    bqkv=torch.empty(3 * H, device=dev, dtype=torch.float32),
                    wo=bf16(H, H),
                    bo=f32(att.out_proj.bias),
                    ln2=(f32(layer.final_layer_norm.weight), f32(layer.final_layer_norm.bias)),
                    w1=bf16(FF, H),
                    b1=f32(layer.feed_forward.intermediate_dense.bias),
                    w2=bf16(H, FF),
                    b2=f32(layer.feed_forward.output_dense.bias),
                )
            )
        self.final_ln = (f32(w.encoder.layer_norm.weight), f32(w.encoder.layer_norm.bias))
        self._pos_w_dgrad: Optional[Tensor] = None
        self._pos_w_dgrad_valid = False
        self.layout_id += 1

    @torch.no_grad()
    def _fill_derived(self) -> None:
        """Operands that are functions of several parameters: fused QKV bias, weight-normed positional conv."""
        w, H = self.weights, self.cfg.hidden_size
        pc = w.encoder.pos_conv_embed.conv
        ops.pack_posconv_weight(pc.parametrizations.weight.original0, pc.parametrizations.weight.original1, dst=self.pos_w_grouped)
        if self.pos_w_grouped is not self.pos_w:
            # [O][tap][group] -> [O][tap][span], the group's columns at its offset inside the super group, zeros elsewhere
            out_channels, group, span = self.pos_w.shape[0], pc.parametrizations.weight.original1.shape[1], self.pos_span
            taps = self.pos_w_grouped.shape[1] // group
            wide = self.pos_w.view(out_channels, taps, span)
            wide.zero_()
            offsets = (torch.arange(out_channels, device=wide.device) % span) // group * group  # first column of each row's group
            columns = offsets[:, None] + torch.arange(group, device=wide.device)[None, :]
            wide.scatter_(2, columns[:, None, :].expand(out_channels, taps, group), self.pos_w_grouped.view(out_channels, taps, group))
        self._pos_w_dgrad_valid = False
        for layer, packed in zip(w.encoder.layers, self.layers):
            att = layer.attention
            for part, linear in enumerate((att.q_proj, att.k_proj, att.v_proj)):
                packed["bqkv"][part * H : (part + 1) * H].copy_(linear.bias.detach())

    @torch.no_grad()
    def _fill(self) -> None:
        w, H = self.weights, self.cfg.hidden_size
        fe = w.feature_extractor.conv_layers
        for index, layer in enumerate(fe[1:], start=1):
            ops.pack_conv_weight(layer.conv.weight, dst=self.conv_w[index])
        ops.cast_bf16(w.feature_projection.projection.weight, dst=self.fp_w)
        for layer, packed in zip(w.encoder.layers, self.layers):
            att = layer.attention
            for part, linear in enumerate((att.q_proj, att.k_proj, att.v_proj)):
                ops.cast_bf16(linear.weight, dst=packed["wqkv"][part * H : (part + 1) * H])
            ops.cast_bf16(att.out_proj.weight, dst=packed["wo"])
            ops.cast_bf16(layer.feed_forward.intermediate_dense.weight, dst=packed["w1"])
            ops.cast_bf16(layer.feed_forward.output_dense.weight, dst=packed["w2"])
        self._fill_derived()

    # -- fused optimiser support ---------------------------------------------------------------
    def shadow_map(self) -> Dict[Tensor, Tensor]:
        """parameter -> bf16 operand (same shape, contiguous) that ``FusedAdam`` keeps up to date itself."""
        self.ensure()
        H = self.cfg.hidden_size
        mapping: Dict[Tensor, Tensor] = {self.weights.feature_projection.projection.weight: self.fp_w}
        for layer, packed in zip(self.weights.encoder.layers, self.layers):
            att = layer.attention
            for part, linear in enumerate((att.q_proj, att.k_proj, att.v_proj)):
                mapping[linear.weight] = packed["wqkv"][part * H : (part + 1) * H]
            mapping[att.out_proj.weight] = packed["wo"]
            mapping[layer.feed_forward.intermediate_dense.weight] = packed["w1"]
            mapping[layer.feed_forward.output_dense.weight] = packed["w2"]
        return mapping

    def after_fused_step(self) -> None:
        """Called by ``FusedAdam`` after a step that wrote every shadowed operand: refresh the derived tensors and
        mark the pack as current (the convolutional feature extractor is frozen or, if it trains, refilled)."""
        if any(p.requires_grad for p in self.weights.feature_extractor.parameters()):
            self._version = None  # conv operands are not shadowed: full refill on next use
            return
        self._fill_derived()
        self._version = self._current_version()

    # -- LayerNorm folded into the q/k/v and the first feed-forward projections (inference launch lists) -------------
    def ensure_folded(self) -> List[Dict[str, Tensor]]:
        """Per layer: ``wqkv`` / ``w1`` with the LayerNorm weight folded in, their column sums and folded biases
        (``ops.fold_layernorm_linear``).  Allocated once per parameter layout and refilled in place whenever the parameters
        change, like the other packed operands (launch lists keep raw pointers)."""
        self.ensure()
        if self._folded is None or self._folded_layout != self.layout_id:
            dev, H, FF = self.device, self.cfg.hidden_size, self.cfg.intermediate_size
            bf16 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.bfloat16)  # noqa: E731
            f32 = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)  # noqa: E731
            self._folded = [
                dict(wqkv=bf16(3 * H, H), sqkv=f32(3 * H), bqkv=f32(3 * H), w1=bf16(FF, H), s1=f32(FF), b1=f32(FF)) for _ in self.layers
            ]
            self._folded_layout = self.layout_id
            self._folded_version = None
        if self._folded_version != self._version:
            H = self.cfg.hidden_size
            with torch.no_grad():
                for layer, folded in zip(self.weights.encoder.layers, self._folded):
                    att = layer.attention
                    g1, b1 = layer.layer_norm.weight, layer.layer_norm.bias
                    for part, linear in enumerate((att.q_proj, att.k_proj, att.v_proj)):
                        rows = slice(part * H, (part + 1) * H)
                        ops.fold_layernorm_linear(linear.weight, linear.bias, g1, b1, folded["wqkv"][rows], folded["sqkv"][rows], folded["bqkv"][rows])
                    dense = layer.feed_forward.intermediate_dense
                    ops.fold_layernorm_linear(dense.weight, dense.bias, layer.final_layer_norm.weight, layer.final_layer_norm.bias, folded["w1"], folded["s1"], folded["b1"])
            self._folded_version = self._version
        return self._folded

    @property
    def pos_w_dgrad(self) -> Tensor:
        """B operand of the positional conv's data-gradient GEMM, packed on first use (training only)."""
        pc = self.weights.encoder.pos_conv_embed.conv
        if self._pos_w_dgrad is None or not self._pos_w_dgrad_valid:
            weight_g, weight_v = pc.parametrizations.weight.original0, pc.parametrizations.weight.original1
            group, span = weight_v.shape[1], self.pos_span
            if span == 64:
                self._pos_w_dgrad = ops.pack_posconv_weight_dgrad(weight_g, weight_v, dst=self._pos_w_dgrad)
            else:
                # [input channel][tap][group-local output channel] -> [input channel][tap][span]: block-diagonal over the super
                # group, exactly as the forward operand (`_fill_derived`)
                grouped = ops.pack_posconv_weight_dgrad(weight_g, weight_v)
                channels, taps = grouped.shape[0], grouped.shape[1] // group
                if self._pos_w_dgrad is None:
                    self._pos_w_dgrad = torch.empty(channels, taps * span, device=grouped.device, dtype=torch.bfloat16)
                wide = self._pos_w_dgrad.view(channels, taps, span)
                wide.zero_()
                offsets = (torch.arange(channels, device=wide.device) % span) // group * group
                columns = offsets[:, None] + torch.arange(group, device=wide.device)[None, :]
                wide.scatter_(2, columns[:, None, :].expand(channels, taps, group), grouped.view(channels, taps, group))
            self._pos_w_dgrad_valid = True
        return self._pos_w_dgrad


Step = Callable[[], None]


@dataclass
class Stochastic:
    """The train()-mode regularisation of ONE forward/backward pair (Hugging Face ``modeling_wav2vec2.py``: feature
    projection dropout 431-433, SpecAugment ``_mask_hidden_states``, encoder dropout 766, LayerDrop 774-777, attention
    dropout, the two hidden dropouts of a layer 742/752).  ``seed`` keys the counter-based masks the forward and the
    backward kernels regenerate (``aph_common.cuh``: ``drop_hash``); LayerDrop is drawn on the host like HF does."""

    seed: int
    hidden_dropout: float = 0.0
    attention_dropout: float = 0.0
    feat_proj_dropout: float = 0.0
    activation_dropout: float = 0.0
    layerdrop: float = 0.0
    mask_time_prob: float = 0.0
    mask_time_length: int = 10
    mask_time_min_masks: int = 2
    mask_feature_prob: float = 0.0
    mask_feature_length: int = 10
    mask_feature_min_masks: int = 0
    skip_layers: Optional[Sequence[bool]] = None  # explicit LayerDrop decisions (tests); None = draw from torch's CPU RNG

    SITE_FEATURE_PROJECTION = 1_000_000
    SITE_ENCODER_INPUT = 1_000_001
    SITE_SPEC_AUGMENT = 1_000_002
    SITE_SPEC_FEATURE = 1_000_003
    SITE_CLASSIFIER_INPUT = 2_000_000  # + hidden-state index (acoustic_model.py:486-488)

    @classmethod
    def from_config(cls, cfg: Wav2Vec2EncoderConfig, seed: int) -> "Stochastic":
        return cls(
            seed, cfg.hidden_dropout, cfg.attention_dropout, cfg.feat_proj_dropout, cfg.activation_dropout, cfg.layerdrop,
            cfg.mask_time_prob if cfg.apply_spec_augment else 0.0, cfg.mask_time_length, cfg.mask_time_min_masks,
            cfg.mask_feature_prob if cfg.apply_spec_augment else 0.0, cfg.mask_feature_length, cfg.mask_feature_min_masks,
        )  # fmt: skip

    def attention(self, layer: int) -> ops.Dropout:
        return ops.Dropout.site(self.attention_dropout, self.seed, 8 * layer)

    def attention_output(self, layer: int) -> ops.Dropout:
        return ops.Dropout.site(self.hidden_dropout, self.seed, 8 * layer + 1)

    def feed_forward_output(self, layer: int) -> ops.Dropout:
        return ops.Dropout.site(self.hidden_dropout, self.seed, 8 * layer + 2)

    def activation(self, layer: int) -> ops.Dropout:
        """``intermediate_dropout`` behind the feed-forward activation (HF:566-569, ``activation_dropout``)."""
        return ops.Dropout.site(self.activation_dropout, self.seed, 8 * layer + 3)

    def feature_projection(self) -> ops.Dropout:
        return ops.Dropout.site(self.feat_proj_dropout, self.seed, self.SITE_FEATURE_PROJECTION)

    def encoder_input(self) -> ops.Dropout:
        return ops.Dropout.site(self.hidden_dropout, self.seed, self.SITE_ENCODER_INPUT)


class EncoderPlan:
    """Workspaces + launch list for one (N, T) shape.  ``run`` enqueues ~200 kernels on the current stream."""

    def __init__(
        self,
        packed: PackedEncoder,
        n_utt: int,
        samples: int,
        ldx: int,
        hidden_blocks: Dict[int, int],
        normalize: bool = True,
        use_lengths: bool = True,
        training: bool = False,
        train_extractor: bool = False,
        arena: Optional[WorkspaceArena] = None,
    ) -> None:
        cfg = packed.cfg
        self.training = training
        self.arena = arena
        self.arena_generation = arena.generation if arena is not None else 0
        self.carver = arena.carver() if arena is not None else None
        self.seq_out: Optional[int] = None  # frames of the unpadded input when the batch was padded up to a bucket
        # the convolutional feature extractor trains too (freeze_feature_encoder = false / UnfreezeSchedule): its
        # pre-LayerNorm conv outputs and activations are kept per layer instead of ping-ponging through two buffers
        self.train_extractor = train_extractor
        if train_extractor and (not training or cfg.feat_extract_norm != "layer"):
            raise NotImplementedError("feature-extractor training is built for training plans of the layer-norm variant (XLS-R, wav2vec2-large-xlsr-53)")
        self.generation = 0  # bumped by every run(): a backward pass checks that its activations are still there
        dev = packed.device
        assert dev is not None
        self.packed = packed
        self.cfg = cfg
        self.n_utt = n_utt
        self.samples = samples
        self.normalize = normalize
        self.use_lengths = use_lengths
        if samples < cfg.conv_kernel[0]:
            raise ValueError("audio is shorter than the first convolution kernel")
        self.conv_lengths = frames_for(samples, cfg)
        if min(self.conv_lengths) < 1:
            raise ValueError(f"{samples} samples are too few for the convolutional feature extractor")
        self.seq = self.conv_lengths[-1]
        self.rows = n_utt * self.seq
        self.ldx = ldx
        self.hidden_blocks = dict(hidden_blocks)  # hidden-state index -> column of X
        H = cfg.hidden_size
        heads = cfg.num_attention_heads
        M = self.rows
        bf16, f32 = torch.bfloat16, torch.float32
        if self.carver is not None:
            z = lambda *shape, dtype=bf16: self.carver.zeros(*shape, device=dev, dtype=dtype)  # noqa: E731
        else:
            z = lambda *shape, dtype=bf16: torch.zeros(*shape, device=dev, dtype=dtype)  # noqa: E731
        self._zeros = z

        self.stats = z(n_utt, 3, dtype=torch.float64)
        self.mean_rstd = z(n_utt, 2, dtype=f32)
        self.frames32 = z(n_utt, dtype=torch.int32)
        self.kernels_dev = torch.tensor(cfg.conv_kernel, device=dev, dtype=torch.int32)
        self.strides_dev = torch.tensor(cfg.conv_stride, device=dev, dtype=torch.int32)
        L = self.conv_lengths
        self.buf_a = z(n_utt * L[0] * 512)
        self.buf_b = z(n_utt * L[1] * 512)
        self.gn_raw = None
        self.gn_stats = None
        if cfg.feat_extract_norm == "group":
            self.gn_raw = z(n_utt * L[0] * 512, dtype=f32)
            self.gn_stats = z(n_utt * 512 * 2, dtype=torch.float64)
        self.fp_in = z(M, 512)
        self.hidden = z(M, H, dtype=f32)
        self.hidden_bf16 = z(M, H)
        self.ln_out = z(M, H)
        self.q = z(n_utt * heads * self.seq * 64)
        self.k = z(n_utt * heads * self.seq * 64)
        self.v = z(n_utt * heads * self.seq * 64)
        self.ctx = z(M, H)
        self.ffn = z(M, cfg.intermediate_size)
        self.x = z(M, ldx)  # classifier feature matrix: [final LN | kept hidden states | dependency probabilities | 0]
        self.row_stats = None if training else z(M, 2 * max(1, H // 256), 2, dtype=f32)  # folded LayerNorms (ops.with_row_stats)
        self.audio_in = None if training else z(n_utt, samples, dtype=f32)  # batches shorter than the bucket are padded into this
        self.captured: Optional[List[Tensor]] = None
        if training:
            # Everything the backward pass reads is kept per layer (sized for 180 GB of HBM: nothing is recomputed
            # except the attention probabilities).  hs[i] = input of layer i (hs[L] = input of the final LayerNorm),
            # mids[i] = residual stream after the attention block of layer i.
            n_layers = len(packed.layers)
            FF = cfg.intermediate_size
            self.hs = [self.hidden] + [z(M, H, dtype=f32) for _ in range(n_layers)]
            self.mids = [z(M, H, dtype=f32) for _ in range(n_layers)]
            self.hidden_fp = z(M, H, dtype=f32)  # feature projection output (input of the positional conv)
            self.pos_pre = z(M, H)               # positional conv + bias, before the GELU
            self.saved = [
                dict(
                    ln1=z(M, H), q=z(n_utt * heads * self.seq * 64), k=z(n_utt * heads * self.seq * 64), v=z(n_utt * heads * self.seq * 64),
                    lse=z(n_utt * heads * self.seq, dtype=f32), ctx=z(M, H), ln2=z(M, H),
                    pre=z(M, FF), act=z(M, FF),
                )
                for _ in range(n_layers)
            ]  # fmt: skip
            # backward workspaces
            self.dh = z(M, H, dtype=f32)
            self.dh_bf16 = z(M, H)
            self.d_ff = z(M, FF)
            self.d_ln = z(M, H, dtype=f32)
            self.d_ctx = z(M, H)
            self.dqkv = z(M, 3 * H)
            self.delta = z(n_utt * heads * self.seq, dtype=f32)
            self.d_fp_in = z(M, 512, dtype=f32)
            # Weight / bias gradients leave the critical path of the backward pass: they run on a second stream, one layer behind
            # at most, and fill the SMs the data-gradient GEMMs leave idle in their last waves (80 tiles on 74 cluster slots at
            # M = 4.9 k).  What they read is therefore double-buffered by layer parity.  APH_BWD_OVERLAP=0 keeps one stream.
            self.bwd_overlap = os.environ.get("APH_BWD_OVERLAP", "1") != "0" and cfg.do_stable_layer_norm
            if self.bwd_overlap:
                self.dh16_ring = [self.dh_bf16, z(M, H), z(M, H), z(M, H)]  # [parity][feed-forward branch, attention branch]
                self.d_ff_ring = [self.d_ff, z(M, FF)]
                self.dqkv_ring = [self.dqkv, z(M, 3 * H)]
                self._side_stream: Optional[torch.cuda.Stream] = None
            if not cfg.do_stable_layer_norm:
                # post-LN ordering: hidden state i is the (normalised) input of layer i; per layer the pre-LayerNorm sums are kept
                self.hs16 = [z(M, H) for _ in range(n_layers)]
                self.e_pre = z(M, H, dtype=f32)  # positional-conv output, before the encoder LayerNorm
                self.mids = []
                self.saved = [
                    dict(q=z(n_utt * heads * self.seq * 64), k=z(n_utt * heads * self.seq * 64), v=z(n_utt * heads * self.seq * 64),
                         lse=z(n_utt * heads * self.seq, dtype=f32), ctx=z(M, H), u=z(M, H, dtype=f32), h1=z(M, H, dtype=f32), h1_16=z(M, H),
                         pre=z(M, FF), act=z(M, FF), v2=z(M, H, dtype=f32))
                    for _ in range(n_layers)
                ]  # fmt: skip
            if train_extractor:
                self.conv_pre = [z(n_utt * length * 512) for length in L]
                self.conv_post = [z(n_utt * length * 512) for length in L]
            self.spec_mask = z(M, dtype=torch.uint8)  # SpecAugment time mask of the last train()-mode run
            self.feature_mask = z(n_utt, H, dtype=torch.uint8)  # ... and its feature-axis mask (mask_feature_prob)
        self.stoch: Optional[Stochastic] = None       # regularisation of the last run (None: eval()-mode arithmetic)
        self.skipped: List[bool] = [False] * len(packed.layers)  # LayerDrop decisions of the last run
        self._layer_spans: List[Tuple[int, int]] = []
        self._drop_gemms: List[Tuple[Any, int, str]] = []  # (GEMM args, layer, "attention_output" | "feed_forward_output" | "activation")

        self._steps: List[Step] = []
        self._build()

    def rebind(self, packed: PackedEncoder) -> None:
        """New packed weights (after an optimizer step) for the same shape: the workspaces stay, the launch
        list (whose GEMM descriptors point at the packed operands) is rebuilt."""
        self.packed = packed
        self._steps = []
        self._layer_spans = []
        self._drop_gemms = []
        self._build()

    # ------------------------------------------------------------------
    def _gemm(self, args: _lib.GemmArgs) -> Step:
        return lambda: ops.run_gemm(args)

    def _build(self) -> None:
        p, cfg = self.packed, self.cfg
        N, L, M, H = self.n_utt, self.conv_lengths, self.rows, cfg.hidden_size
        eps = cfg.layer_norm_eps
        steps = self._steps
        layer_norm = cfg.feat_extract_norm == "layer"

        # conv layers 1..6: implicit GEMM on the channels-last activation (+ LayerNorm + GELU in place)
        src, dst = self.buf_a, self.buf_b
        for i in range(1, len(L)):
            kernel, stride = cfg.conv_kernel[i], cfg.conv_stride[i]
            if self.train_extractor:
                src, dst = self.conv_post[i - 1], self.conv_pre[i]
            args = ops.make_gemm_args(
                src,
                p.conv_w[i],
                a_rows=L[i],
                a_inner=kernel * 512,
                a_row_stride=stride * 512,
                batch=N,
                a_batch_stride=L[i - 1] * 512,
                bias=p.conv_bias[i],
                gelu=not layer_norm,
                out_bf16=dst,
                ld_bf16=512,
                out_batch_rows=L[i],
            )
            steps.append(self._gemm(args))
            if layer_norm:
                g, b = p.conv_ln[i]
                activated = self.conv_post[i] if self.train_extractor else dst
                steps.append(
                    lambda dst=dst, rows=N * L[i], g=g, b=b, activated=activated: ops.layernorm_rows(
                        dst, rows, 512, 512, g, b, 1e-5, gelu=True, out_bf16=activated, ld_bf16=512
                    )
                )
            src, dst = dst, src
        conv_out = self.conv_post[-1] if self.train_extractor else src  # [N, T', 512]
        self.conv_out = conv_out

        # feature projection: LN -> Linear, padded frames zeroed (HF:753-756), fp32 residual stream + bf16 copy
        g, b = p.fp_ln
        steps.append(lambda: ops.layernorm_rows(conv_out, M, 512, 512, g, b, eps, out_bf16=self.fp_in, ld_bf16=512))
        steps.append(
            self._gemm(
                ops.make_gemm_args(
                    self.fp_in,
                    p.fp_w,
                    a_rows=M,
                    a_inner=512,
                    a_row_stride=512,
                    bias=p.fp_b,
                    out_f32=self.hidden_fp if self.training else self.hidden,
                    ld_f32=H,
                    out_bf16=self.hidden_bf16,
                    ld_bf16=H,
                    lengths=self.frames32 if self.use_lengths else None,
                    len_period=self.seq,
                )
            )
        )
        if self.training:
            steps.append(self._regularise_projection)
        # positional conv embedding: hidden += gelu(grouped_conv(hidden)) (HF:764-765, 353-368)
        taps = cfg.num_conv_pos_embeddings
        pos_args = ops.make_gemm_args(
            self.hidden_bf16,
            p.pos_w,
            a_rows=self.seq,
            a_inner=H,
            a_row_stride=H,
            batch=N,
            a_batch_stride=self.seq * H,
            mode=_lib.APH_GEMM_TAPS,
            tap_pad=taps // 2,
            n=H,
            k=taps * p.pos_span,
            bias=p.pos_b,
            gelu=True,
            resid=self.hidden_fp if self.training else self.hidden,
            ld_resid=H,
            out_f32=self.hidden,
            ld_f32=H,
            out_batch_rows=self.seq,
            aux_bf16=self.pos_pre if self.training else None,
            ld_aux=H,
        )
        pos_args.taps_span = p.pos_span
        steps.append(self._gemm(pos_args))
        if self.training:
            steps.append(self._regularise_encoder_input)
        heads = cfg.num_attention_heads
        FF = cfg.intermediate_size
        if not cfg.do_stable_layer_norm:
            self._build_post_ln(steps, heads, FF)
            return
        # Inference option (APH_FOLD_LN=1): the two LayerNorms of a layer folded into the GEMMs around them.  out-proj / FFN2 leave a
        # bf16 copy of the residual stream and per-row (sum, sum of squares); the q/k/v and FFN1 projections read that copy, a
        # gamma-folded weight and apply mean / rstd in their epilogue (ops.with_row_stats / with_layernorm).  Only the first
        # LayerNorm of layer 0 (its input comes from the positional conv) and the encoder's final LayerNorm still run as kernels:
        # 146 launches instead of 193.  Measured (profiles/r02_layernorm_fold.md): the 47 LayerNorm launches it removes (0.94 ms)
        # are paid back almost entirely inside the producing GEMMs (third staging tile -> 4 instead of 5 pipeline stages, +14 us per
        # FFN2), the step gains 1.2 % and the GEMM's own roofline figure drops by 7 points because its time now includes the
        # LayerNorm work — so it is off by default.
        self.fold_ln = (not self.training) and os.environ.get("APH_FOLD_LN", "0") == "1" and H % 256 == 0
        folded = p.ensure_folded() if self.fold_ln else None
        for index, lw in enumerate(p.layers):
            if self.training:
                sv = self.saved[index]
                h_in, h_mid, h_out = self.hs[index], self.mids[index], self.hs[index + 1]
                ln1, ln2, q, k, v, ctx, ffn = sv["ln1"], sv["ln2"], sv["q"], sv["k"], sv["v"], sv["ctx"], sv["act"]
                lse, pre = sv["lse"], sv["pre"]
            else:
                h_in = h_mid = h_out = self.hidden
                ln1 = ln2 = self.ln_out
                q, k, v, ctx, ffn = self.q, self.k, self.v, self.ctx, self.ffn
                lse = pre = None
            steps.append(lambda index=index, h_in=h_in: self._keep_hidden(index, h_in))
            span_start = len(steps)
            g1, b1 = lw["ln1"]
            if self.fold_ln and index > 0:
                qkv_args = ops.make_qkv_args(self.hidden_bf16, folded[index]["wqkv"], folded[index]["bqkv"], q, k, v, rows=M, seq=self.seq, heads=heads)
                steps.append(self._gemm(ops.with_layernorm(qkv_args, self.row_stats, folded[index]["sqkv"], H, eps)))
            else:
                steps.append(lambda g1=g1, b1=b1, h_in=h_in, ln1=ln1: ops.layernorm_rows(h_in, M, H, H, g1, b1, eps, out_bf16=ln1, ld_bf16=H))
                steps.append(
                    self._gemm(ops.make_qkv_args(ln1, lw["wqkv"], lw["bqkv"], q, k, v, rows=M, seq=self.seq, heads=heads))
                )
            steps.append(
                lambda q=q, k=k, v=v, ctx=ctx, lse=lse, index=index: ops.attention(
                    q, k, v, ctx, self.att_lengths, N, heads, self.seq, lse, self.stoch.attention(index) if self.stoch else ops.NO_DROPOUT
                )
            )
            out_args = ops.make_gemm_args(ctx, lw["wo"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["bo"], resid=h_in, ld_resid=H, out_f32=h_mid, ld_f32=H)
            if self.fold_ln:
                out_args.out_bf16, out_args.ld_bf16 = self.hidden_bf16.data_ptr(), H
                ops.with_row_stats(out_args, self.row_stats)
            steps.append(self._gemm(out_args))
            g2, b2 = lw["ln2"]
            if self.fold_ln:
                inner_args = ops.make_gemm_args(
                    self.hidden_bf16, folded[index]["w1"], a_rows=M, a_inner=H, a_row_stride=H, bias=folded[index]["b1"], gelu=True, out_bf16=ffn, ld_bf16=FF
                )
                ops.with_layernorm(inner_args, self.row_stats, folded[index]["s1"], H, eps)
            else:
                steps.append(lambda g2=g2, b2=b2, h_mid=h_mid, ln2=ln2: ops.layernorm_rows(h_mid, M, H, H, g2, b2, eps, out_bf16=ln2, ld_bf16=H))
                inner_args = ops.make_gemm_args(
                    ln2, lw["w1"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["b1"], gelu=True, out_bf16=ffn, ld_bf16=FF, aux_bf16=pre, ld_aux=FF
                )
            steps.append(self._gemm(inner_args))
            ffn_args = ops.make_gemm_args(ffn, lw["w2"], a_rows=M, a_inner=FF, a_row_stride=FF, bias=lw["b2"], resid=h_mid, ld_resid=H, out_f32=h_out, ld_f32=H)
            if self.fold_ln and index + 1 < len(p.layers):
                ffn_args.out_bf16, ffn_args.ld_bf16 = self.hidden_bf16.data_ptr(), H
                ops.with_row_stats(ffn_args, self.row_stats)
            steps.append(self._gemm(ffn_args))
            self._layer_spans.append((span_start, len(steps)))
            if self.training:
                self._drop_gemms += [(out_args, index, "attention_output"), (ffn_args, index, "feed_forward_output"), (inner_args, index, "activation")]
        self.h_last = self.hs[len(p.layers)] if self.training else self.hidden
        gf, bf = p.final_ln
        steps.append(lambda: ops.layernorm_rows(self.h_last, M, H, H, gf, bf, eps, out_bf16=self.x, ld_bf16=self.ldx))
        steps.append(lambda: self._keep_hidden(len(p.layers), self.h_last))

    def _build_post_ln(self, steps: List[Step], heads: int, FF: int) -> None:
        """The post-LN encoder ordering of wav2vec2-base style checkpoints (``do_stable_layer_norm = False``; HF
        ``Wav2Vec2Encoder`` / ``Wav2Vec2EncoderLayer``): the encoder LayerNorm follows the positional convolution, every layer is
        ``h = LN(h + attention(h)); h = final_LN(h + FFN(h))`` and hidden state ``i`` is the input of layer ``i``.  Inference
        only: the fp32 stream is normalised in place and the same kernel writes the bf16 GEMM operand."""
        p, cfg = self.packed, self.cfg
        N, M, H, eps = self.n_utt, self.rows, cfg.hidden_size, cfg.layer_norm_eps
        if self.training:
            self._build_post_ln_training(steps, heads, FF)
            return
        hidden, ln16 = self.hidden, self.ln_out
        n_layers = len(p.layers)

        def normalise(gamma: Tensor, beta: Tensor, target: Tensor, ld_target: int) -> None:
            """fp32 stream normalised in place + the bf16 copy the next GEMM (or the classifiers) read"""
            if H in (512, 1024):
                ops.layernorm_rows(hidden, M, H, H, gamma, beta, eps, out_f32=hidden, ld_f32=H, out_bf16=target, ld_bf16=ld_target)
            else:
                ops.layernorm_any(hidden, H, M, H, gamma, beta, eps, out_f32=hidden, ld_f32=H, out_bf16=target, ld_bf16=ld_target)

        ge, be = p.final_ln  # encoder.layer_norm
        first_target, first_ld = (self.x, self.ldx) if n_layers == 0 else (ln16, H)
        steps.append(lambda: normalise(ge, be, first_target, first_ld))
        for index, lw in enumerate(p.layers):
            steps.append(lambda index=index: self._keep_hidden(index, hidden))
            span_start = len(steps)
            steps.append(self._gemm(ops.make_qkv_args(ln16, lw["wqkv"], lw["bqkv"], self.q, self.k, self.v, rows=M, seq=self.seq, heads=heads)))
            steps.append(lambda: ops.attention(self.q, self.k, self.v, self.ctx, self.att_lengths, N, heads, self.seq))
            steps.append(self._gemm(ops.make_gemm_args(self.ctx, lw["wo"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["bo"], resid=hidden, ld_resid=H, out_f32=hidden, ld_f32=H)))
            g1, b1 = lw["ln1"]
            steps.append(lambda g1=g1, b1=b1: normalise(g1, b1, ln16, H))
            steps.append(self._gemm(ops.make_gemm_args(ln16, lw["w1"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["b1"], gelu=True, out_bf16=self.ffn, ld_bf16=FF)))
            steps.append(self._gemm(ops.make_gemm_args(self.ffn, lw["w2"], a_rows=M, a_inner=FF, a_row_stride=FF, bias=lw["b2"], resid=hidden, ld_resid=H, out_f32=hidden, ld_f32=H)))
            g2, b2 = lw["ln2"]
            last = index == n_layers - 1
            target, ld_target = (self.x, self.ldx) if last else (ln16, H)  # the last hidden state is the classifier input OUTPUT
            steps.append(lambda g2=g2, b2=b2, target=target, ld_target=ld_target: normalise(g2, b2, target, ld_target))
            self._layer_spans.append((span_start, len(steps)))
        self.h_last = hidden
        steps.append(lambda: self._keep_hidden(n_layers, hidden))

    def _build_post_ln_training(self, steps: List[Step], heads: int, FF: int) -> None:
        """Post-LN ordering with everything the backward pass reads kept per layer (``_backward_post_ln_layers``)."""
        p, cfg = self.packed, self.cfg
        N, M, H, eps = self.n_utt, self.rows, cfg.hidden_size, cfg.layer_norm_eps
        n_layers = len(p.layers)
        hs, hs16 = self.hs, self.hs16
        ge, be = p.final_ln  # encoder.layer_norm

        def first_target() -> Tuple[Tensor, int]:
            return (self.x, self.ldx) if n_layers == 0 else (hs16[0], H)

        def encoder_norm() -> None:
            self.e_pre.copy_(self.hidden)  # hs[0] is self.hidden: keep the pre-LayerNorm value for the backward pass
            target, ld = first_target()
            ops.layernorm_rows(self.hidden, M, H, H, ge, be, eps, out_f32=hs[0], ld_f32=H, out_bf16=target, ld_bf16=ld)
            st = self.stoch
            if st is not None and st.encoder_input().threshold:  # HF: hidden_states = self.dropout(self.layer_norm(hidden_states))
                ops.dropout_2d(hs[0], H, M, H, st.encoder_input(), out_f32=hs[0], ld_f32=H, out_bf16=target, ld_bf16=ld)

        steps.append(encoder_norm)
        for index, lw in enumerate(p.layers):
            sv = self.saved[index]
            h_in, h_in16 = hs[index], hs16[index]
            last = index == n_layers - 1
            steps.append(lambda index=index, h_in=h_in: self._keep_hidden(index, h_in))
            span_start = len(steps)
            steps.append(self._gemm(ops.make_qkv_args(h_in16, lw["wqkv"], lw["bqkv"], sv["q"], sv["k"], sv["v"], rows=M, seq=self.seq, heads=heads)))
            steps.append(
                lambda sv=sv, index=index: ops.attention(
                    sv["q"], sv["k"], sv["v"], sv["ctx"], self.att_lengths, N, heads, self.seq, sv["lse"],
                    self.stoch.attention(index) if self.stoch else ops.NO_DROPOUT,
                )
            )
            out_args = ops.make_gemm_args(sv["ctx"], lw["wo"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["bo"], resid=h_in, ld_resid=H, out_f32=sv["u"], ld_f32=H)
            steps.append(self._gemm(out_args))
            g1, b1 = lw["ln1"]
            steps.append(lambda sv=sv, g1=g1, b1=b1: ops.layernorm_rows(sv["u"], M, H, H, g1, b1, eps, out_f32=sv["h1"], ld_f32=H, out_bf16=sv["h1_16"], ld_bf16=H))
            inner_args = ops.make_gemm_args(sv["h1_16"], lw["w1"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["b1"], gelu=True, out_bf16=sv["act"], ld_bf16=FF,
                                            aux_bf16=sv["pre"], ld_aux=FF)  # fmt: skip
            steps.append(self._gemm(inner_args))
            ffn_args = ops.make_gemm_args(sv["act"], lw["w2"], a_rows=M, a_inner=FF, a_row_stride=FF, bias=lw["b2"], resid=sv["h1"], ld_resid=H, out_f32=sv["v2"], ld_f32=H)
            steps.append(self._gemm(ffn_args))
            g2, b2 = lw["ln2"]
            target, ld_target = (self.x, self.ldx) if last else (hs16[index + 1], H)
            steps.append(lambda sv=sv, g2=g2, b2=b2, index=index, target=target, ld_target=ld_target: ops.layernorm_rows(
                sv["v2"], M, H, H, g2, b2, eps, out_f32=hs[index + 1], ld_f32=H, out_bf16=target, ld_bf16=ld_target))  # fmt: skip
            self._layer_spans.append((span_start, len(steps)))
            self._drop_gemms += [(out_args, index, "attention_output"), (ffn_args, index, "feed_forward_output"), (inner_args, index, "activation")]
        self.h_last = hs[n_layers]
        steps.append(lambda: self._keep_hidden(n_layers, hs[n_layers]))

    def _backward_post_ln_layers(self, d_x: Tensor, need_encoder: bool, group: Any, done: Any, wgrad: Any, branch_gradient: Any) -> None:
        """Backward pass of the post-LN layers: on return ``self.dh`` is the gradient of hidden state 0 (the encoder LayerNorm's
        output).  Per layer ``u = h + Wo attn(h); h1 = LN1(u); v2 = h1 + W2 gelu(W1 h1); out = LN2(v2)``."""
        p, cfg, st = self.packed, self.cfg, self.stoch
        N, M, H, FF, heads, seq, eps = self.n_utt, self.rows, cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, self.seq, cfg.layer_norm_eps
        dh, dh16 = self.dh, self.dh_bf16
        n_layers = len(p.layers)
        dh.copy_(d_x[:, :H])  # the last hidden state IS the classifier input OUTPUT (no final LayerNorm in this ordering)
        for index in reversed(range(n_layers)):
            lw, sv = p.layers[index], self.saved[index]
            column = self.hidden_blocks.get(index + 1)
            if column is not None and index + 1 < n_layers:
                ops.add_2d(dh, H, d_x[:, column:], self.ldx, M, H)
            if need_encoder:
                flat, g = group(
                    [
                        ("feed_forward.output_dense.weight", (H, FF)), ("feed_forward.output_dense.bias", (H,)),
                        ("feed_forward.intermediate_dense.weight", (FF, H)), ("feed_forward.intermediate_dense.bias", (FF,)),
                        ("final_layer_norm.weight", (H,)), ("final_layer_norm.bias", (H,)),
                        ("attention.out_proj.weight", (H, H)), ("attention.out_proj.bias", (H,)),
                        ("attention.qkv.weight", (3 * H, H)), ("attention.qkv.bias", (3 * H,)),
                        ("layer_norm.weight", (H,)), ("layer_norm.bias", (H,)),
                    ]
                )  # fmt: skip
            else:
                flat, g = None, {}

            def finish_group() -> None:
                if need_encoder:
                    wqkv, bqkv = g.pop("attention.qkv.weight"), g.pop("attention.qkv.bias")
                    for name, weight_part, bias_part in zip(("q_proj", "k_proj", "v_proj"), wqkv.split(H), bqkv.split(H)):
                        g[f"attention.{name}.weight"] = weight_part
                        g[f"attention.{name}.bias"] = bias_part
                    done(flat, g, f"encoder.layers.{index}.")

            if self.skipped[index]:
                if need_encoder:
                    flat.zero_()
                finish_group()
                continue
            # out = LN2(v2)
            g2, _ = lw["ln2"]
            ops.layernorm_backward(sv["v2"], H, dh, H, M, H, g2, eps, None, 0, dh, H, g.get("final_layer_norm.weight"), g.get("final_layer_norm.bias"))
            # v2 = h1 + dropout(W2 dropout(gelu(W1 h1 + b1)) + b2): dh is d(v2)
            dropped = branch_gradient(st.feed_forward_output(index) if st else ops.NO_DROPOUT)
            args = ops.make_dgrad_args(dh16, lw["w2"], rows=M, ld_dy=H, k=H, n=FF, ld_w=FF, gelu_bwd=sv["pre"], ld_gelu_bwd=FF, out_bf16=self.d_ff, ld_bf16=FF)
            inner = st.activation(index) if st else ops.NO_DROPOUT
            args.drop_threshold, args.drop_seed, args.drop_scale = inner.threshold, inner.seed, inner.scale
            ops.run_gemm(args)
            if need_encoder:
                wgrad(g["feed_forward.output_dense.weight"], dh16, H, H, sv["act"], FF, FF)
                if dropped:
                    ops.colsum_bf16(dh16, M, H, H, out=g["feed_forward.output_dense.bias"])
                else:
                    ops.colsum_f32(dh, M, H, H, out=g["feed_forward.output_dense.bias"])
                wgrad(g["feed_forward.intermediate_dense.weight"], self.d_ff, FF, FF, sv["h1_16"], H, H)
                ops.colsum_bf16(self.d_ff, M, FF, FF, out=g["feed_forward.intermediate_dense.bias"])
            # d(h1) = d(v2) + d_ff W1 (residual epilogue), then h1 = LN1(u)
            ops.run_gemm(ops.make_dgrad_args(self.d_ff, lw["w1"], rows=M, ld_dy=FF, k=FF, n=H, ld_w=H, resid=dh, ld_resid=H, out_f32=self.d_ln, ld_f32=H))
            g1, _ = lw["ln1"]
            ops.layernorm_backward(sv["u"], H, self.d_ln, H, M, H, g1, eps, None, 0, dh, H, g.get("layer_norm.weight"), g.get("layer_norm.bias"))
            # u = h + dropout(Wo attn(Wqkv h) + bo): dh is d(u)
            dropped = branch_gradient(st.attention_output(index) if st else ops.NO_DROPOUT)
            ops.run_gemm(ops.make_dgrad_args(dh16, lw["wo"], rows=M, ld_dy=H, k=H, n=H, ld_w=H, out_bf16=self.d_ctx, ld_bf16=H))
            if need_encoder:
                wgrad(g["attention.out_proj.weight"], dh16, H, H, sv["ctx"], H, H)
                if dropped:
                    ops.colsum_bf16(dh16, M, H, H, out=g["attention.out_proj.bias"])
                else:
                    ops.colsum_f32(dh, M, H, H, out=g["attention.out_proj.bias"])
            ops.attention_backward(
                sv["q"], sv["k"], sv["v"], sv["ctx"], self.d_ctx, sv["lse"], self.delta, self.dqkv, self.att_lengths, N, heads, seq,
                st.attention(index) if st else ops.NO_DROPOUT,
            )  # fmt: skip
            if need_encoder:
                wgrad(g["attention.qkv.weight"], self.dqkv, 3 * H, 3 * H, self.hs16[index], H, H)
                ops.colsum_bf16(self.dqkv, M, 3 * H, 3 * H, out=g["attention.qkv.bias"])
            # d(h) = d(u) + dqkv Wqkv, in place through the residual epilogue
            ops.run_gemm(ops.make_dgrad_args(self.dqkv, lw["wqkv"], rows=M, ld_dy=3 * H, k=3 * H, n=H, ld_w=H, resid=dh, ld_resid=H, out_f32=dh, ld_f32=H))
            finish_group()

    def _regularise_projection(self) -> None:
        """train() mode: dropout of the feature projection output (HF:431-433), then SpecAugment (HF
        ``_mask_hidden_states``: masked frames <- ``masked_spec_embed``); fp32 in place + the bf16 copy the positional
        convolution reads.  Padded frames are zero already and stay zero."""
        st = self.stoch
        if st is None:
            return
        H, M = self.cfg.hidden_size, self.rows
        embed = getattr(self.packed.weights, "masked_spec_embed", None)
        mask = None
        if st.mask_time_prob > 0.0 and embed is not None:
            ops.spec_augment_mask(
                self.att_lengths, self.seq, st.mask_time_prob, st.mask_time_length, st.mask_time_min_masks,
                ops.mix_seed(st.seed, Stochastic.SITE_SPEC_AUGMENT), self.spec_mask,
            )  # fmt: skip
            mask = self.spec_mask
        drop = st.feature_projection()
        if mask is not None or drop.threshold != 0:
            fill = embed.detach().float().contiguous() if mask is not None else None
            ops.dropout_2d(self.hidden_fp, H, M, H, drop, out_f32=self.hidden_fp, ld_f32=H, out_bf16=self.hidden_bf16, ld_bf16=H, row_mask=mask, row_fill=fill)
        self.spec_active = mask is not None
        self.feature_active = st.mask_feature_prob > 0.0
        if self.feature_active:  # HF _mask_hidden_states: spans along the hidden axis are zeroed for every frame of an utterance
            full = torch.full((self.n_utt,), H, device=self.hidden_fp.device, dtype=torch.int32)
            ops.spec_augment_mask(full, H, st.mask_feature_prob, st.mask_feature_length, st.mask_feature_min_masks,
                                  ops.mix_seed(st.seed, Stochastic.SITE_SPEC_FEATURE), self.feature_mask)  # fmt: skip
            ops.mask_columns(self.hidden_fp, H, self.n_utt, self.seq, H, self.feature_mask, self.hidden_bf16, H)

    def _regularise_encoder_input(self) -> None:
        """train() mode: ``hidden_states = self.dropout(hidden_states + position_embeddings)`` (HF:765-766)."""
        st = self.stoch
        if st is None or not self.cfg.do_stable_layer_norm:  # post-LN: the dropout follows the encoder LayerNorm (_build_post_ln)
            return
        drop = st.encoder_input()
        if drop.threshold:
            H = self.cfg.hidden_size
            ops.dropout_2d(self.hidden, H, self.rows, H, drop, out_f32=self.hidden, ld_f32=H)

    def _keep_hidden(self, index: int, hidden: Tensor) -> None:
        """Hidden state ``index`` of HF's ``hidden_states`` tuple is live in ``hidden`` right now
        (for the last index: the final LayerNorm output, already in X)."""
        last = len(self.packed.layers)
        if self.captured is not None:
            if index < last:
                self.captured.append(hidden.clone())
            else:
                self.captured.append(self.x[:, : self.cfg.hidden_size].float())
        column = self.hidden_blocks.get(index)
        if column is not None and index < last:
            ops.cast_bf16_2d(hidden, self.cfg.hidden_size, self.x[:, column:], self.ldx, self.rows, self.cfg.hidden_size)

    # ------------------------------------------------------------------
    def run(self, audio: Tensor, lengths: Tensor, frames64: Tensor, capture: bool = False, stochastic: Optional[Stochastic] = None) -> None:
        """Enqueues the whole encoder.  ``audio`` fp32 [N, T] and ``lengths`` int64 [N] on the device;
        ``frames64`` (int64 [N], caller-owned) receives the per-utterance frame counts.  ``stochastic`` (training
        plans only) switches the train()-mode regularisation on for this run and the backward pass that follows it."""
        p, cfg = self.packed, self.cfg
        N = self.n_utt
        if getattr(self, "fold_ln", False):
            p.ensure_folded()  # refilled in place when the parameters changed since the last run
        if self.arena is not None:
            if self.arena_generation != self.arena.generation:
                raise RuntimeError("this launch list was carved from a workspace arena that has been replaced")
            # other launch lists overlay the same memory: the zero padding columns of X must be zero again (everything else a run
            # reads it has written itself; the attention kernel clears the context rows of query tiles it skips)
            self.x.zero_()
        self.captured = [] if capture else None
        self.generation += 1
        if stochastic is not None and not self.training:
            raise RuntimeError("train()-mode regularisation needs a training plan")
        self.stoch = stochastic
        self.spec_active = False
        self.feature_active = False
        n_layers = len(p.layers)
        self.skipped = [False] * n_layers
        if stochastic is not None:
            if stochastic.skip_layers is not None:
                self.skipped = [bool(flag) for flag in stochastic.skip_layers]
            elif stochastic.layerdrop > 0.0:
                self.skipped = (torch.rand(n_layers) < stochastic.layerdrop).tolist()  # HF:774-777, host RNG like HF
        for args, layer, site in self._drop_gemms:
            drop = ops.NO_DROPOUT if stochastic is None else getattr(stochastic, site)(layer)
            args.drop_threshold, args.drop_seed, args.drop_scale = drop.threshold, drop.seed, drop.scale
        ops.frame_lengths(lengths, self.kernels_dev, self.strides_dev, self.frames32, frames64)
        if self.use_lengths:
            self.att_lengths = self.frames32
        else:
            self.att_lengths = torch.full_like(self.frames32, self.seq)
        mean_rstd = None
        if self.normalize:
            ops.wave_stats(audio, lengths, self.stats, self.mean_rstd)
            mean_rstd = self.mean_rstd
        if self.train_extractor:
            g, b = p.conv_ln[0]
            self._wave = (audio, lengths, mean_rstd)  # the backward pass of the first convolution reads the waveform again
            ops.conv0_raw(audio, lengths, mean_rstd, p.conv0_w, p.conv_bias[0], self.conv_pre[0])
            ops.layernorm_rows(self.conv_pre[0], N * self.conv_lengths[0], 512, 512, g, b, 1e-5, gelu=True, out_bf16=self.conv_post[0], ld_bf16=512)
        elif cfg.feat_extract_norm == "layer":
            g, b = p.conv_ln[0]
            ops.conv0_ln_gelu(audio, lengths, mean_rstd, p.conv0_w, p.conv_bias[0], g, b, 1e-5, self.buf_a, self.use_lengths)
        else:
            g, b = p.conv_ln[0]
            ops.conv0_gn_gelu(audio, lengths, mean_rstd, p.conv0_w, p.conv_bias[0], g, b, 1e-5, self.gn_raw, self.gn_stats, self.buf_a)
        position = 0
        for layer, (start, end) in enumerate(self._layer_spans):
            for step in self._steps[position:start]:
                step()
            if self.skipped[layer]:
                self.hs[layer + 1].copy_(self.hs[layer])  # LayerDrop: the layer is the identity for this batch
                if not cfg.do_stable_layer_norm:  # post-LN: the bf16 operand of the next layer (or the classifier input) follows
                    if layer + 1 < n_layers:
                        self.hs16[layer + 1].copy_(self.hs16[layer])
                    else:
                        ops.cast_bf16_2d(self.hs[layer], cfg.hidden_size, self.x, self.ldx, self.rows, cfg.hidden_size)
            else:
                for step in self._steps[start:end]:
                    step()
            position = end
        for step in self._steps[position:]:
            step()

    # ------------------------------------------------------------------ backward (training plans only)
    @torch.no_grad()
    def backward(
        self,
        d_x: Tensor,
        need_encoder: bool,
        need_projection: bool,
        on_group_ready: Optional[Callable[[Tensor, Dict[str, Tensor]], None]] = None,
        need_extractor: bool = False,
    ) -> Dict[str, Tensor]:
        """Backward pass of everything ``run`` enqueued, from the gradient of the classifier feature matrix.

        ``d_x`` fp32 ``[M, ldx]`` is dL/dX (X = ``[final LayerNorm | kept hidden states | ...]``).  Returns
        fp32 gradients keyed by the Hugging Face parameter names of ``Wav2Vec2Weights`` for the transformer
        (``need_encoder``), the feature projection (``need_projection``) and — on plans built with
        ``train_extractor=True`` — the convolutional feature extractor (``need_extractor``; frozen in the reference's
        default configuration, ``default_config.toml:40``, trained after its ``UnfreezeSchedule`` fires).

        Every Linear contributes two tcgen05 GEMMs that read the forward pass's own buffers:
        ``dX = dY W`` (W as an MN-major B operand) and ``dW = dY^T X`` (both operands MN-major, split-K
        with TMA reduce-add stores); GELU' is fused into the FFN2 data-gradient epilogue and residual
        accumulation into the LayerNorm backward kernel.

        The gradients of one layer live in ONE flat buffer (the returned tensors are views of it);
        ``on_group_ready(flat, views)`` is called as soon as the kernels producing a group are enqueued, so a
        data-parallel caller can start that group's all-reduce while earlier layers are still computing."""
        if not self.training:
            raise RuntimeError("this EncoderPlan was built for inference: no activations were kept")
        p, cfg = self.packed, self.cfg
        w = p.weights
        N, M, H, FF = self.n_utt, self.rows, cfg.hidden_size, cfg.intermediate_size
        heads, eps, seq = cfg.num_attention_heads, cfg.layer_norm_eps, self.seq
        dev = d_x.device
        grads: Dict[str, Tensor] = {}
        n_layers = len(p.layers)
        dh, dh16 = self.dh, self.dh_bf16
        st = self.stoch  # train()-mode regularisation of the forward pass this backward pass belongs to

        def branch_gradient(drop: ops.Dropout, target: Optional[Tensor] = None) -> bool:
            """target (dh16) <- bf16(dh o keep * scale): gradient of a residual branch behind the forward's epilogue dropout."""
            target = dh16 if target is None else target
            if drop.threshold:
                ops.dropout_2d(dh, H, M, H, drop, out_bf16=target, ld_bf16=H)
                return True
            ops.cast_bf16_2d(dh, H, target, H, M, H)
            return False

        def group(shapes: Sequence[Tuple[str, Tuple[int, ...]]]) -> Tuple[Tensor, Dict[str, Tensor]]:
            # one allocation, one split, one view per matrix: the host enqueues the backward pass barely ahead of the GPU, and
            # slicing + viewing every gradient separately was 2 ms of Python per step
            sizes = [math.prod(shape) for _, shape in shapes]
            flat = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
            parts = flat.split_with_sizes(sizes)
            views = {name: (part if len(shape) == 1 else part.view(shape)) for (name, shape), part in zip(shapes, parts)}
            return flat, views

        def done(flat: Tensor, views: Dict[str, Tensor], prefix: str) -> None:
            named = {prefix + name: value for name, value in views.items()}
            grads.update(named)
            if on_group_ready is not None:
                on_group_ready(flat, named)

        def wgrad(out: Tensor, dy: Tensor, ld_dy: int, m: int, x: Tensor, ld_x: int, n: int) -> None:
            ops.run_gemm(ops.make_wgrad_args(dy, x, out, rows=M, m=m, ld_dy=ld_dy, n=n, ld_x=ld_x, ld_out=n))

        post_ln = not cfg.do_stable_layer_norm
        gf, _ = p.final_ln
        if post_ln:
            self._backward_post_ln_layers(d_x, need_encoder, group, done, wgrad, branch_gradient)
        else:
            # final LayerNorm (HF:792): X[:, :H] = LN(hs[L])
            flat, g = group([("weight", (H,)), ("bias", (H,))]) if need_encoder else (None, {})
            ops.layernorm_backward(self.hs[n_layers], H, d_x, self.ldx, M, H, gf, eps, None, 0, dh, H, g.get("weight"), g.get("bias"))
            if need_encoder:
                done(flat, g, "encoder.layer_norm.")

        # ---- the layers, last to first.  With `bwd_overlap` the weight and bias gradients of a layer are enqueued on a second
        # stream behind two events of the main stream (d_ff ready; dqkv ready) and handed out (`done`) one layer later, when
        # the main stream has waited for them — which is also what allows it to overwrite the buffers they read.
        overlap = bool(getattr(self, "bwd_overlap", False)) and _BACKWARD_OVERLAP and need_encoder and not post_ln
        side: Optional[torch.cuda.Stream] = None
        if overlap:
            if self._side_stream is None or self._side_stream.device != dev:
                self._side_stream = torch.cuda.Stream(device=dev)
            side = self._side_stream
        main = torch.cuda.current_stream(dev) if overlap else None
        pending: List[Tuple[Any, Tensor, Dict[str, Tensor], str]] = []  # (side-stream event, flat, views, prefix) of layers not handed out yet

        def hand_out(keep: int) -> None:
            while len(pending) > keep:
                event, flat_done, views_done, prefix_done = pending.pop(0)
                main.wait_event(event)
                done(flat_done, views_done, prefix_done)

        for index in (() if post_ln else reversed(range(n_layers))):
            lw, sv = p.layers[index], self.saved[index]
            if need_encoder:
                flat, g = group(
                    [
                        ("feed_forward.output_dense.weight", (H, FF)), ("feed_forward.output_dense.bias", (H,)),
                        ("feed_forward.intermediate_dense.weight", (FF, H)), ("feed_forward.intermediate_dense.bias", (FF,)),
                        ("final_layer_norm.weight", (H,)), ("final_layer_norm.bias", (H,)),
                        ("attention.out_proj.weight", (H, H)), ("attention.out_proj.bias", (H,)),
                        ("attention.qkv.weight", (3 * H, H)), ("attention.qkv.bias", (3 * H,)),
                        ("layer_norm.weight", (H,)), ("layer_norm.bias", (H,)),
                    ]
                )  # fmt: skip
            else:
                flat, g = None, {}
            column = self.hidden_blocks.get(index + 1)
            if column is not None and index + 1 < n_layers:  # hidden state index+1 is also a classifier input (OUTPUT_i)
                ops.add_2d(dh, H, d_x[:, column:], self.ldx, M, H)
            if self.skipped[index]:  # LayerDrop: identity in the forward pass, no gradient for its parameters
                if need_encoder:
                    flat.zero_()
                    wqkv, bqkv = g.pop("attention.qkv.weight"), g.pop("attention.qkv.bias")
                    for part, name in enumerate(("q_proj", "k_proj", "v_proj")):
                        g[f"attention.{name}.weight"] = wqkv[part * H : (part + 1) * H]
                        g[f"attention.{name}.bias"] = bqkv[part * H : (part + 1) * H]
                    if overlap:  # keep the hand-out order (the all-reduce of a data-parallel caller is issued per group, in this order)
                        event = torch.cuda.Event()
                        event.record(main)
                        pending.append((event, flat, g, f"encoder.layers.{index}."))
                    else:
                        done(flat, g, f"encoder.layers.{index}.")
                continue
            parity = index & 1
            dh16_ff = self.dh16_ring[2 * parity] if overlap else dh16
            dh16_att = self.dh16_ring[2 * parity + 1] if overlap else dh16
            d_ff = self.d_ff_ring[parity] if overlap else self.d_ff
            dqkv = self.dqkv_ring[parity] if overlap else self.dqkv
            if overlap:
                hand_out(1)  # the layer of the same parity (two back) has been waited for: its operand buffers are free again
            # ---- feed forward: h_out = h_mid + dropout(W2 gelu(W1 LN2(h_mid) + b1) + b2)
            dropped_ff = branch_gradient(st.feed_forward_output(index) if st else ops.NO_DROPOUT, dh16_ff)
            if need_encoder and not dropped_ff:  # (reads the fp32 stream the LayerNorm backward below rewrites: stays on this stream)
                ops.colsum_f32(dh, M, H, H, out=g["feed_forward.output_dense.bias"])
            args = ops.make_dgrad_args(dh16_ff, lw["w2"], rows=M, ld_dy=H, k=H, n=FF, ld_w=FF, gelu_bwd=sv["pre"], ld_gelu_bwd=FF,
                                       out_bf16=d_ff, ld_bf16=FF)  # fmt: skip
            inner = st.activation(index) if st else ops.NO_DROPOUT  # the mask of the dropout behind the activation
            args.drop_threshold, args.drop_seed, args.drop_scale = inner.threshold, inner.seed, inner.scale
            ops.run_gemm(args)

            def feed_forward_parameters() -> None:
                wgrad(g["feed_forward.output_dense.weight"], dh16_ff, H, H, sv["act"], FF, FF)
                if dropped_ff:
                    ops.colsum_bf16(dh16_ff, M, H, H, out=g["feed_forward.output_dense.bias"])
                wgrad(g["feed_forward.intermediate_dense.weight"], d_ff, FF, FF, sv["ln2"], H, H)
                ops.colsum_bf16(d_ff, M, FF, FF, out=g["feed_forward.intermediate_dense.bias"])

            if need_encoder:
                if overlap:
                    ready = torch.cuda.Event()
                    ready.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(ready)
                        feed_forward_parameters()
                else:
                    feed_forward_parameters()
            ops.run_gemm(ops.make_dgrad_args(d_ff, lw["w1"], rows=M, ld_dy=FF, k=FF, n=H, ld_w=H, out_f32=self.d_ln, ld_f32=H))
            g2, _ = lw["ln2"]
            ops.layernorm_backward(self.mids[index], H, self.d_ln, H, M, H, g2, eps, dh, H, dh, H,
                                   g.get("final_layer_norm.weight"), g.get("final_layer_norm.bias"))  # fmt: skip
            # ---- attention block: h_mid = h_in + dropout(Wo attention(Wqkv LN1(h_in)) + bo)
            dropped_att = branch_gradient(st.attention_output(index) if st else ops.NO_DROPOUT, dh16_att)
            if need_encoder and not dropped_att:
                ops.colsum_f32(dh, M, H, H, out=g["attention.out_proj.bias"])
            ops.run_gemm(ops.make_dgrad_args(dh16_att, lw["wo"], rows=M, ld_dy=H, k=H, n=H, ld_w=H, out_bf16=self.d_ctx, ld_bf16=H))
            ops.attention_backward(
                sv["q"], sv["k"], sv["v"], sv["ctx"], self.d_ctx, sv["lse"], self.delta, dqkv, self.att_lengths, N, heads, seq,
                st.attention(index) if st else ops.NO_DROPOUT,
            )  # fmt: skip

            def attention_parameters() -> None:
                wgrad(g["attention.out_proj.weight"], dh16_att, H, H, sv["ctx"], H, H)
                if dropped_att:
                    ops.colsum_bf16(dh16_att, M, H, H, out=g["attention.out_proj.bias"])
                wgrad(g["attention.qkv.weight"], dqkv, 3 * H, 3 * H, sv["ln1"], H, H)
                ops.colsum_bf16(dqkv, M, 3 * H, 3 * H, out=g["attention.qkv.bias"])

            finished = None
            if need_encoder:
                if overlap:
                    ready = torch.cuda.Event()
                    ready.record(main)
                    with torch.cuda.stream(side):
                        side.wait_event(ready)
                        attention_parameters()
                        finished = torch.cuda.Event()
                        finished.record(side)
                else:
                    attention_parameters()
            ops.run_gemm(ops.make_dgrad_args(dqkv, lw["wqkv"], rows=M, ld_dy=3 * H, k=3 * H, n=H, ld_w=H, out_f32=self.d_ln, ld_f32=H))
            g1, _ = lw["ln1"]
            ops.layernorm_backward(self.hs[index], H, self.d_ln, H, M, H, g1, eps, dh, H, dh, H, g.get("layer_norm.weight"), g.get("layer_norm.bias"))
            if need_encoder:
                # the fused QKV gradient is handed out as its three nn.Linear slices (views of the same buffer)
                wqkv, bqkv = g.pop("attention.qkv.weight"), g.pop("attention.qkv.bias")
                for name, weight_part, bias_part in zip(("q_proj", "k_proj", "v_proj"), wqkv.split(H), bqkv.split(H)):
                    g[f"attention.{name}.weight"] = weight_part
                    g[f"attention.{name}.bias"] = bias_part
                if overlap:
                    # the LayerNorm gradients of this group are written by the main stream: the hand-out also follows this point
                    pending.append((finished, flat, g, f"encoder.layers.{index}."))
                else:
                    done(flat, g, f"encoder.layers.{index}.")
        if overlap:
            hand_out(0)

        column = self.hidden_blocks.get(0)
        if column is not None and n_layers > 0:
            ops.add_2d(dh, H, d_x[:, column:], self.ldx, M, H)
        # ---- positional conv embedding (HF:764-766, 353-368): hs[0] = dropout(h_fp + gelu(conv(h_fp) + b))
        if st is not None and st.encoder_input().threshold:
            ops.dropout_2d(dh, H, M, H, st.encoder_input(), out_f32=dh, ld_f32=H)
        if post_ln:  # hidden state 0 = dropout(LN_enc(h_fp + gelu(conv(h_fp)))): through the encoder LayerNorm
            flat, g = group([("weight", (H,)), ("bias", (H,))]) if need_encoder else (None, {})
            ops.layernorm_backward(self.e_pre, H, dh, H, M, H, gf, eps, None, 0, dh, H, g.get("weight"), g.get("bias"))
            if need_encoder:
                done(flat, g, "encoder.layer_norm.")
        taps = cfg.num_conv_pos_embeddings
        pc = w.encoder.pos_conv_embed.conv
        weight_g, weight_v = pc.parametrizations.weight.original0, pc.parametrizations.weight.original1
        ops.gelu_backward_bf16(dh, H, self.pos_pre, H, M, H, dh16, H)  # dh16 = d(conv output)
        if need_encoder:
            flat, g = group(
                [("bias", (H,)), ("parametrizations.weight.original0", tuple(weight_g.shape)), ("parametrizations.weight.original1", tuple(weight_v.shape))]
            )
            ops.colsum_bf16(dh16, M, H, H, out=g["bias"])
            if p.pos_span == 64:
                raw = torch.empty(taps, H, 256, device=dev, dtype=torch.float32)
                args = ops.make_wgrad_args(dh16, self.hidden_bf16, raw, rows=seq, m=H, ld_dy=H, n=H, ld_x=H, ld_out=256)
                args.mode, args.n_taps, args.tap_pad = _lib.APH_GEMM_DIAG_TAPS, taps, taps // 2
                args.k_batch, args.a_batch_stride, args.b_seg_stride = N, seq * H, seq * H
                args.out_batch_rows = H
                ops.run_gemm(args)
                block_width = 256
            else:
                # groups that do not tile 256 channels (wav2vec2-base: 16 x 48): one full [H][H] weight-gradient GEMM per tap, the
                # input shifted by tap - taps/2 frames inside each utterance (zero fill outside); the group-diagonal entries are
                # read out below.  ~H/group times the necessary FLOPs — a coverage path, not a tuned one
                raw = torch.empty(taps, H, H, device=dev, dtype=torch.float32)
                for tap in range(taps):
                    args = ops.make_wgrad_args(dh16, self.hidden_bf16, raw[tap], rows=seq, m=H, ld_dy=H, n=H, ld_x=H, ld_out=H)
                    args.k_batch, args.a_batch_stride, args.b_seg_stride = N, seq * H, seq * H
                    args.b_k_shift = tap - taps // 2
                    ops.run_gemm(args)
                block_width = H
            ops.posconv_weight_backward(raw, weight_g, weight_v, g["parametrizations.weight.original0"], g["parametrizations.weight.original1"],
                                        block_width=block_width)  # fmt: skip
            done(flat, g, "encoder.pos_conv_embed.conv.")
        embed = getattr(w, "masked_spec_embed", None)
        need_embed = st is not None and self.spec_active and embed is not None and embed.requires_grad
        if need_extractor and not self.train_extractor:
            raise RuntimeError("this EncoderPlan did not keep the feature extractor's activations (train_extractor=False)")
        if need_projection or need_embed or need_extractor:
            # data gradient of the grouped conv: the same sliding-tap GEMM with flipped taps, accumulated onto dh
            dgrad_args = ops.make_gemm_args(
                dh16, p.pos_w_dgrad, a_rows=seq, a_inner=H, a_row_stride=H, batch=N, a_batch_stride=seq * H, mode=_lib.APH_GEMM_TAPS,
                tap_pad=taps // 2 - 1, n=H, k=taps * p.pos_span, resid=dh, ld_resid=H, out_f32=dh, ld_f32=H, out_batch_rows=seq,
            )  # fmt: skip
            dgrad_args.taps_span = p.pos_span
            ops.run_gemm(dgrad_args)
            if st is not None:
                # SpecAugment: masked frames were replaced by masked_spec_embed (its gradient: their sum), then the
                # feature projection dropout; both are functions of (seed, row, column) only
                if self.feature_active:
                    ops.mask_columns(dh, H, N, seq, H, self.feature_mask)
                if self.spec_active:
                    flat, g = group([("masked_spec_embed", (H,))])
                    ops.masked_rows_backward(dh, H, M, H, self.spec_mask, g["masked_spec_embed"])
                    if need_embed:
                        done(flat, g, "")
                if st.feature_projection().threshold:
                    ops.dropout_2d(dh, H, M, H, st.feature_projection(), out_f32=dh, ld_f32=H)
        if need_projection or need_extractor:
            # ---- feature projection (HF:422-434) behind the padded-frame zeroing (HF:753-756)
            if self.use_lengths:
                ops.mask_rows(dh, H, M, H, self.frames32, seq)
            ops.cast_bf16_2d(dh, H, dh16, H, M, H)
            flat, g = group([("projection.weight", (H, 512)), ("projection.bias", (H,)), ("layer_norm.weight", (512,)), ("layer_norm.bias", (512,))])
            wgrad(g["projection.weight"], dh16, H, H, self.fp_in, 512, 512)
            ops.colsum_f32(dh, M, H, H, out=g["projection.bias"])
            ops.run_gemm(ops.make_dgrad_args(dh16, p.fp_w, rows=M, ld_dy=H, k=H, n=512, ld_w=512, out_f32=self.d_fp_in, ld_f32=512))
            gp, _ = p.fp_ln
            ops.layernorm_backward(self.conv_out, 512, self.d_fp_in, 512, M, 512, gp, eps, None, 0, self.d_fp_in, 512, g["layer_norm.weight"], g["layer_norm.bias"])
            done(flat, g, "feature_projection.")
        if need_extractor:
            self._backward_extractor(self.d_fp_in, group, done)
        return grads

    def _backward_extractor(self, d_post: Tensor, group: Any, done: Any) -> None:
        """Backward pass of the convolutional feature extractor (HF:275-323, 382-419, layer-norm variant) from ``d_post`` =
        gradient of its output (fp32 ``[N * T', 512]``).  Per layer: LayerNorm + GELU backward from the kept pre-LayerNorm conv
        output (one fused kernel, also the conv bias gradient), weight gradient as a GEMM over overlapping-row windows of the
        kept input activation (one K segment per utterance), data gradient GEMM + col2im; the first convolution reads the
        normalised waveform again."""
        p, cfg = self.packed, self.cfg
        N, L = self.n_utt, self.conv_lengths
        dev = d_post.device
        audio, lengths, mean_rstd = self._wave
        layers = p.weights.feature_extractor.conv_layers
        for i in reversed(range(len(L))):
            kernel, stride = cfg.conv_kernel[i], cfg.conv_stride[i]
            has_bias = p.conv_bias[i] is not None
            shapes = [("conv.weight", tuple(layers[i].conv.weight.shape)), ("layer_norm.weight", (512,)), ("layer_norm.bias", (512,))]
            if has_bias:
                shapes.append(("conv.bias", (512,)))
            flat, g = group(shapes)
            gamma, beta = p.conv_ln[i]
            rows = N * L[i]
            d_pre = torch.empty(rows, 512, device=dev, dtype=torch.bfloat16)
            ops.ln_gelu_backward_512(self.conv_pre[i], d_post, 512, rows, gamma, beta, 1e-5, d_pre, g["layer_norm.weight"], g["layer_norm.bias"], g.get("conv.bias"))
            if i == 0:
                ops.conv0_weight_backward(d_pre, audio, lengths, mean_rstd, g["conv.weight"].view(512, kernel))
            else:
                width = kernel * 512
                raw = torch.empty(512, width, device=dev, dtype=torch.float32)
                args = ops.make_wgrad_args(d_pre, self.conv_post[i - 1], raw, rows=L[i], m=512, ld_dy=512, n=width, ld_x=stride * 512, ld_out=width)
                args.k_batch, args.a_batch_stride, args.b_seg_stride = N, L[i] * 512, L[i - 1] * 512
                ops.run_gemm(args)
                g["conv.weight"].copy_(raw.view(512, kernel, 512).permute(0, 2, 1))  # [O][k][C] -> Conv1d's [O][C][k]
                d_cols = torch.empty(rows, width, device=dev, dtype=torch.float32)
                ops.run_gemm(ops.make_dgrad_args(d_pre, p.conv_w[i], rows=rows, ld_dy=512, k=512, n=width, ld_w=width, out_f32=d_cols, ld_f32=width))
                d_in = torch.empty(N * L[i - 1], 512, device=dev, dtype=torch.float32)
                full = torch.full((N,), L[i - 1], device=dev, dtype=torch.int32)
                ops.conv_input_backward(d_cols, full, N, L[i - 1], 512, L[i], kernel, stride, 0, 0, False, d_in, 512)
                d_post = d_in
            done(flat, g, f"feature_extractor.conv_layers.{i}.")
