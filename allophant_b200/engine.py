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

from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib, ops
from .network.wav2vec2 import Wav2Vec2EncoderConfig, Wav2Vec2Weights


def conv_out_length(length: int, kernel: int, stride: int) -> int:
    return (length - kernel) // stride + 1


def frames_for(samples: int, cfg: Wav2Vec2EncoderConfig) -> List[int]:
    """Output length of every conv layer for ``samples`` input samples."""
    lengths = []
    for kernel, stride in zip(cfg.conv_kernel, cfg.conv_stride):
        samples = conv_out_length(samples, kernel, stride)
        lengths.append(samples)
    return lengths


class PackedEncoder:
    """bf16 GEMM operands packed from the fp32 master parameters (re-packed when they change)."""

    def __init__(self, weights: Wav2Vec2Weights) -> None:
        self.weights = weights
        self.cfg = weights.config
        self._version: Optional[Tuple[int, ...]] = None
        self.device: Optional[torch.device] = None

    def _current_version(self) -> Tuple[int, ...]:
        def version(p: Tensor) -> int:
            try:
                return p._version
            except RuntimeError:  # inference tensors do not track a version counter
                return -1

        params = list(self.weights.parameters())
        return tuple(version(p) for p in params) + tuple(p.data_ptr() for p in params)

    def ensure(self) -> None:
        version = self._current_version()
        if version != self._version:
            self._pack()
            self._version = version

    @torch.no_grad()
    def _pack(self) -> None:
        w = self.weights
        cfg = self.cfg
        first = next(w.parameters())
        if not first.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the model to a GPU (`model.to('cuda')`)")
        self.device = first.device
        if cfg.hidden_size % 256 != 0 or cfg.head_dim != 64:
            raise NotImplementedError("the CUDA encoder supports head_dim 64 and hidden sizes that are multiples of 256")
        if any(c != 512 for c in cfg.conv_dim) or cfg.conv_kernel[0] != 10 or cfg.conv_stride[0] != 5:
            raise NotImplementedError("the CUDA feature extractor supports the wav2vec2 layout (7 x 512 channels, k0=10, s0=5)")

        def f32(p: Tensor) -> Tensor:
            return p.detach().float().contiguous()

        fe = w.feature_extractor.conv_layers
        self.conv0_w = f32(fe[0].conv.weight).view(512, 10)
        self.conv_bias = [f32(l.conv.bias) if l.conv.bias is not None else None for l in fe]
        self.conv_ln = [(f32(l.layer_norm.weight), f32(l.layer_norm.bias)) if hasattr(l, "layer_norm") else None for l in fe]
        self.conv_w = [None] + [ops.pack_conv_weight(l.conv.weight) for l in fe[1:]]
        fp = w.feature_projection
        self.fp_ln = (f32(fp.layer_norm.weight), f32(fp.layer_norm.bias))
        self.fp_w = ops.cast_bf16(fp.projection.weight)
        self.fp_b = f32(fp.projection.bias)
        pc = w.encoder.pos_conv_embed.conv
        self.pos_w = ops.pack_posconv_weight(pc.parametrizations.weight.original0, pc.parametrizations.weight.original1)
        self.pos_b = f32(pc.bias)
        self.layers = []
        for layer in w.encoder.layers:
            att = layer.attention
            wqkv = torch.cat([att.q_proj.weight.detach(), att.k_proj.weight.detach(), att.v_proj.weight.detach()], 0)
            bqkv = torch.cat([att.q_proj.bias.detach(), att.k_proj.bias.detach(), att.v_proj.bias.detach()], 0)
            self.layers.append(
                dict(
                    ln1=(f32(layer.layer_norm.weight), f32(layer.layer_norm.bias)),
                    wqkv=ops.cast_bf16(wqkv),
                    bqkv=f32(bqkv),
                    wo=ops.cast_bf16(att.out_proj.weight),
                    bo=f32(att.out_proj.bias),
                    ln2=(f32(layer.final_layer_norm.weight), f32(layer.final_layer_norm.bias)),
                    w1=ops.cast_bf16(layer.feed_forward.intermediate_dense.weight),
                    b1=f32(layer.feed_forward.intermediate_dense.bias),
                    w2=ops.cast_bf16(layer.feed_forward.output_dense.weight),
                    b2=f32(layer.feed_forward.output_dense.bias),
                )
            )
        self.final_ln = (f32(w.encoder.layer_norm.weight), f32(w.encoder.layer_norm.bias))


Step = Callable[[], None]


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
    ) -> None:
        cfg = packed.cfg
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
        z = lambda *shape, dtype=bf16: torch.zeros(*shape, device=dev, dtype=dtype)  # noqa: E731

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
        self.t_v = (self.seq + 7) // 8 * 8
        self.q = z(n_utt * heads * self.seq * 64)
        self.k = z(n_utt * heads * self.seq * 64)
        self.vt = z(n_utt * heads * 64 * self.t_v)
        self.ctx = z(M, H)
        self.ffn = z(M, cfg.intermediate_size)
        self.x = z(M, ldx)  # classifier feature matrix: [final LN | kept hidden states | dependency probabilities | 0]
        self.captured: Optional[List[Tensor]] = None

        self._steps: List[Step] = []
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
                steps.append(
                    lambda dst=dst, rows=N * L[i], g=g, b=b: ops.layernorm_rows(
                        dst, rows, 512, 512, g, b, 1e-5, gelu=True, out_bf16=dst, ld_bf16=512
                    )
                )
            src, dst = dst, src
        conv_out = src  # [N, T', 512]

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
                    out_f32=self.hidden,
                    ld_f32=H,
                    out_bf16=self.hidden_bf16,
                    ld_bf16=H,
                    lengths=self.frames32 if self.use_lengths else None,
                    len_period=self.seq,
                )
            )
        )
        # positional conv embedding: hidden += gelu(grouped_conv(hidden)) (HF:764-765, 353-368)
        taps = cfg.num_conv_pos_embeddings
        if H // cfg.num_conv_pos_embedding_groups != 64:
            raise NotImplementedError("positional conv groups must be 64 channels wide")
        steps.append(
            self._gemm(
                ops.make_gemm_args(
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
                    k=taps * 64,
                    bias=p.pos_b,
                    gelu=True,
                    resid=self.hidden,
                    ld_resid=H,
                    out_f32=self.hidden,
                    ld_f32=H,
                    out_batch_rows=self.seq,
                )
            )
        )
        if not cfg.do_stable_layer_norm:
            raise NotImplementedError("post-LN wav2vec2 encoders (do_stable_layer_norm=False) are not implemented yet")

        heads = cfg.num_attention_heads
        for index, lw in enumerate(p.layers):
            steps.append(lambda index=index: self._keep_hidden(index))
            g1, b1 = lw["ln1"]
            steps.append(lambda g1=g1, b1=b1: ops.layernorm_rows(self.hidden, M, H, H, g1, b1, eps, out_bf16=self.ln_out, ld_bf16=H))
            steps.append(
                self._gemm(
                    ops.make_qkv_args(self.ln_out, lw["wqkv"], lw["bqkv"], self.q, self.k, self.vt, rows=M, seq=self.seq, heads=heads, t_v=self.t_v)
                )
            )
            steps.append(lambda: ops.attention(self.q, self.k, self.vt, self.ctx, self.att_lengths, N, heads, self.seq, self.t_v))
            steps.append(
                self._gemm(
                    ops.make_gemm_args(
                        self.ctx, lw["wo"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["bo"], resid=self.hidden, ld_resid=H, out_f32=self.hidden, ld_f32=H
                    )
                )
            )
            g2, b2 = lw["ln2"]
            steps.append(lambda g2=g2, b2=b2: ops.layernorm_rows(self.hidden, M, H, H, g2, b2, eps, out_bf16=self.ln_out, ld_bf16=H))
            steps.append(
                self._gemm(
                    ops.make_gemm_args(
                        self.ln_out, lw["w1"], a_rows=M, a_inner=H, a_row_stride=H, bias=lw["b1"], gelu=True, out_bf16=self.ffn, ld_bf16=cfg.intermediate_size
                    )
                )
            )
            steps.append(
                self._gemm(
                    ops.make_gemm_args(
                        self.ffn,
                        lw["w2"],
                        a_rows=M,
                        a_inner=cfg.intermediate_size,
                        a_row_stride=cfg.intermediate_size,
                        bias=lw["b2"],
                        resid=self.hidden,
                        ld_resid=H,
                        out_f32=self.hidden,
                        ld_f32=H,
                    )
                )
            )
        gf, bf = p.final_ln
        steps.append(lambda: ops.layernorm_rows(self.hidden, M, H, H, gf, bf, eps, out_bf16=self.x, ld_bf16=self.ldx))
        steps.append(lambda: self._keep_hidden(len(p.layers)))

    def _keep_hidden(self, index: int) -> None:
        """Hidden state ``index`` of HF's ``hidden_states`` tuple is live in ``self.hidden`` right now
        (for the last index: the final LayerNorm output, already in X)."""
        last = len(self.packed.layers)
        if self.captured is not None:
            if index < last:
                self.captured.append(self.hidden.clone())
            else:
                self.captured.append(self.x[:, : self.cfg.hidden_size].float())
        column = self.hidden_blocks.get(index)
        if column is not None and index < last:
            ops.cast_bf16_2d(self.hidden, self.cfg.hidden_size, self.x[:, column:], self.ldx, self.rows, self.cfg.hidden_size)

    # ------------------------------------------------------------------
    def run(self, audio: Tensor, lengths: Tensor, frames64: Tensor, capture: bool = False) -> None:
        """Enqueues the whole encoder.  ``audio`` fp32 [N, T] and ``lengths`` int64 [N] on the device;
        ``frames64`` (int64 [N], caller-owned) receives the per-utterance frame counts."""
        p, cfg = self.packed, self.cfg
        N = self.n_utt
        self.captured = [] if capture else None
        ops.frame_lengths(lengths, self.kernels_dev, self.strides_dev, self.frames32, frames64)
        if self.use_lengths:
            self.att_lengths = self.frames32
        else:
            self.att_lengths = torch.full_like(self.frames32, self.seq)
        mean_rstd = None
        if self.normalize:
            ops.wave_stats(audio, lengths, self.stats, self.mean_rstd)
            mean_rstd = self.mean_rstd
        if cfg.feat_extract_norm == "layer":
            g, b = p.conv_ln[0]
            ops.conv0_ln_gelu(audio, lengths, mean_rstd, p.conv0_w, p.conv_bias[0], g, b, 1e-5, self.buf_a, self.use_lengths)
        else:
            g, b = p.conv_ln[0]
            ops.conv0_gn_gelu(audio, lengths, mean_rstd, p.conv0_w, p.conv_bias[0], g, b, 1e-5, self.gn_raw, self.gn_stats, self.buf_a)
        for step in self._steps:
            step()
