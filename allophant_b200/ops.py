"""Tensor-level wrappers over the C ABI (``include/allophant_b200.h``).

Every function takes CUDA tensors, launches on ``torch.cuda.current_stream()`` and
returns without synchronising.  torch is used for device memory and streams only;
the arithmetic happens in ``liballophant_b200.so``.  Non-CUDA tensors are rejected:
there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import CtcHead, GemmArgs, HeadBlock, check, lib


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> ctypes.c_void_p:
    """``cudaStream_t`` of torch's current stream (raw accessors: ~1 us instead of ~12 us per launch)."""
    if _raw_stream is not None and _raw_device is not None:
        return ctypes.c_void_p(_raw_stream(_raw_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "allophant_b200 kernels run on CUDA tensors only (no CPU fallback exists); "
                f"got a tensor on {t.device}"
            )


def launch_count() -> int:
    return int(lib.aph_launch_count())


def reset_launch_count() -> None:
    lib.aph_reset_launch_count()


def set_pdl(enabled: bool) -> bool:
    """Programmatic dependent launch between the library's kernels (on by default; ``APH_PDL=0`` disables it at start-up).
    Returns the previous setting."""
    return bool(lib.aph_set_pdl(1 if enabled else 0))


def set_gemm_tail_split(enabled: bool) -> bool:
    """Tail split of the tcgen05 GEMM (see ``aph_set_gemm_tail_split`` in ``include/allophant_b200.h``; on by default,
    ``APH_GEMM_TAIL_SPLIT=0`` disables it at start-up).  Returns the previous setting."""
    return bool(lib.aph_set_gemm_tail_split(1 if enabled else 0))


def set_attention_kernel(mode: int) -> int:
    """Which attention forward kernel runs: 0 = chosen by problem size (default), 1 = 64-key blocks with two CTAs per SM, 2 = the
    persistent query-tile-pair kernel (``aph_set_attention_kernel``).  Returns the previous setting."""
    return int(lib.aph_set_attention_kernel(int(mode)))


# --------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------
def make_gemm_args(
    a: Tensor,
    w: Tensor,
    *,
    a_rows: int,
    a_inner: int,
    a_row_stride: int,
    batch: int = 1,
    a_batch_stride: int = 0,
    k: Optional[int] = None,
    n: Optional[int] = None,
    mode: int = _lib.APH_GEMM_ROWS,
    tap_pad: int = 0,
    bias: Optional[Tensor] = None,
    gelu: bool = False,
    scale: float = 1.0,
    resid: Optional[Tensor] = None,
    ld_resid: int = 0,
    out_f32: Optional[Tensor] = None,
    ld_f32: int = 0,
    out_bf16: Optional[Tensor] = None,
    ld_bf16: int = 0,
    out_batch_rows: int = 0,
    lengths: Optional[Tensor] = None,
    len_period: int = 0,
    aux_bf16: Optional[Tensor] = None,
    ld_aux: int = 0,
    gelu_bwd: Optional[Tensor] = None,
    ld_gelu_bwd: int = 0,
) -> GemmArgs:
    """Builds an ``aph_gemm_args`` (store epilogue).  The caller keeps the tensors alive."""
    _require_cuda(a, w, bias, resid, out_f32, out_bf16, lengths, aux_bf16, gelu_bwd)
    g = GemmArgs()
    g.a = a.data_ptr()
    g.a_row_stride, g.a_batch_stride = a_row_stride, a_batch_stride
    g.a_rows, g.a_inner, g.batch = a_rows, a_inner, batch
    g.mode, g.tap_pad = mode, tap_pad
    g.b = w.data_ptr()
    g.n = w.shape[0] if n is None else n
    g.k = w.shape[1] if k is None else k
    g.epilogue = _lib.APH_EPI_STORE
    g.gelu = int(gelu)
    g.scale = scale
    g.bias = _ptr(bias)
    g.resid, g.ld_resid = _ptr(resid), ld_resid
    g.out_f32, g.ld_f32 = _ptr(out_f32), ld_f32
    g.out_bf16, g.ld_bf16 = _ptr(out_bf16), ld_bf16
    g.out_batch_rows = out_batch_rows
    g.lengths, g.len_period = _ptr(lengths), len_period
    g.aux_bf16, g.ld_aux = _ptr(aux_bf16), ld_aux
    g.gelu_bwd, g.ld_gelu_bwd = _ptr(gelu_bwd), ld_gelu_bwd
    return g


def make_dgrad_args(
    dy: Tensor,
    w: Tensor,
    *,
    rows: int,
    ld_dy: int,
    k: int,
    n: int,
    ld_w: int,
    scale: float = 1.0,
    gelu_bwd: Optional[Tensor] = None,
    ld_gelu_bwd: int = 0,
    resid: Optional[Tensor] = None,
    ld_resid: int = 0,
    out_f32: Optional[Tensor] = None,
    ld_f32: int = 0,
    out_bf16: Optional[Tensor] = None,
    ld_bf16: int = 0,
) -> GemmArgs:
    """Data gradient of ``y = x @ w.T``: ``dx[rows, n] = dy[rows, k] @ w[k, n]`` with ``w`` read in its
    forward layout ``[out = k][in = n]`` as an MN-major B operand (no transposed copy)."""
    _require_cuda(dy, w, gelu_bwd, resid, out_f32, out_bf16)
    g = GemmArgs()
    g.a = dy.data_ptr()
    g.a_row_stride, g.a_batch_stride = ld_dy, 0
    g.a_rows, g.a_inner, g.batch = rows, k, 1
    g.mode = _lib.APH_GEMM_ROWS
    g.b, g.n, g.k = w.data_ptr(), n, k
    g.b_mn_major, g.b_row_stride, g.k_seq, g.k_batch = 1, ld_w, k, 1
    g.epilogue = _lib.APH_EPI_STORE
    g.scale = scale
    g.gelu_bwd, g.ld_gelu_bwd = _ptr(gelu_bwd), ld_gelu_bwd
    g.resid, g.ld_resid = _ptr(resid), ld_resid
    g.out_f32, g.ld_f32 = _ptr(out_f32), ld_f32
    g.out_bf16, g.ld_bf16 = _ptr(out_bf16), ld_bf16
    return g


def make_wgrad_args(
    dy: Tensor,
    x: Tensor,
    out_f32: Tensor,
    *,
    rows: int,
    m: int,
    ld_dy: int,
    n: int,
    ld_x: int,
    ld_out: int,
    scale: float = 1.0,
) -> GemmArgs:
    """Weight gradient of ``y = x @ w.T``: ``dw[m, n] = sum_r dy[r, m] * x[r, n]`` with both operands read
    frame-major as they were produced (MN-major A and B); the frame axis may have any length."""
    _require_cuda(dy, x, out_f32)
    g = GemmArgs()
    g.a = dy.data_ptr()
    g.a_row_stride, g.a_batch_stride = ld_dy, 0
    g.a_rows, g.a_inner, g.batch = m, m, 1
    g.mode = _lib.APH_GEMM_ROWS
    g.b, g.n, g.k = x.data_ptr(), n, 0
    g.a_mn_major, g.b_mn_major, g.b_row_stride, g.k_seq, g.k_batch = 1, 1, ld_x, rows, 1
    g.epilogue = _lib.APH_EPI_STORE
    g.scale = scale
    g.out_f32, g.ld_f32 = out_f32.data_ptr(), ld_out
    return g


def make_qkv_args(a: Tensor, w_qkv: Tensor, bias_qkv: Tensor, q: Tensor, k: Tensor, v: Tensor, *, rows: int, seq: int, heads: int) -> GemmArgs:
    """QKV projection with the scatter epilogue: Q (pre-scaled), K and V as bf16 ``[n_utt*heads, seq, 64]``."""
    _require_cuda(a, w_qkv, bias_qkv, q, k, v)
    hidden = heads * 64
    g = GemmArgs()
    g.a = a.data_ptr()
    g.a_row_stride, g.a_batch_stride = a.stride(0), 0
    g.a_rows, g.a_inner, g.batch = rows, hidden, 1
    g.mode = _lib.APH_GEMM_ROWS
    g.b, g.n, g.k = w_qkv.data_ptr(), 3 * hidden, hidden
    g.epilogue = _lib.APH_EPI_QKV
    g.scale = 1.0
    g.bias = bias_qkv.data_ptr()
    g.len_period = seq
    g.q, g.kmat, g.vmat = q.data_ptr(), k.data_ptr(), v.data_ptr()
    g.heads, g.q_scale = heads, 0.125 * 1.4426950408889634  # head_dim^-0.5 * log2(e)
    return g


def run_gemm(args: GemmArgs) -> None:
    check(lib.aph_gemm_bf16(ctypes.byref(args), _stream()), "aph_gemm_bf16")


def fold_layernorm_linear(weight: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor, out_weight: Tensor, out_colsum: Tensor, out_bias: Tensor) -> None:
    """LayerNorm folded into the Linear behind it: ``out_weight = bf16(weight * gamma)``, ``out_colsum = out_weight.sum(1)``,
    ``out_bias = bias + weight @ beta`` (all written in place; ``out_weight`` may be a row slice of a larger operand)."""
    _require_cuda(weight, bias, gamma, beta, out_weight, out_colsum, out_bias)
    w = weight.detach().float().contiguous()
    n, k = w.shape
    if not out_weight.is_contiguous() or out_weight.shape != (n, k) or out_weight.dtype != torch.bfloat16:
        raise ValueError("out_weight must be a contiguous bf16 [n, k] tensor")
    b = None if bias is None else bias.detach().float().contiguous()
    check(
        lib.aph_fold_layernorm_linear(
            w.data_ptr(), _ptr(b), gamma.detach().float().contiguous().data_ptr(), beta.detach().float().contiguous().data_ptr(), n, k,
            out_weight.data_ptr(), out_colsum.data_ptr(), out_bias.data_ptr(), _stream(),
        ),
        "aph_fold_layernorm_linear",
    )  # fmt: skip


def with_row_stats(args: GemmArgs, stats: Tensor) -> GemmArgs:
    """Producer side of a folded LayerNorm: the GEMM (fp32 output + residual + bf16 copy) also leaves per-row partial sums and
    sums of squares in ``stats`` fp32 ``[rows, 2 * ceil(n / 256), 2]``."""
    _require_cuda(stats)
    args.row_stats = stats.data_ptr()
    args.row_stats_slots = stats.shape[1]
    return args


def with_layernorm(args: GemmArgs, stats: Tensor, colsum: Tensor, cols: int, eps: float) -> GemmArgs:
    """Consumer side: ``args.a`` holds the UN-normalised rows (bf16), ``args.b`` the gamma-folded weight, ``args.bias`` the folded
    bias; the epilogue applies the row statistics (``fold_layernorm_linear``, ``with_row_stats``)."""
    _require_cuda(stats, colsum)
    args.ln_stats = stats.data_ptr()
    args.ln_slots = stats.shape[1]
    args.ln_cols = cols
    args.ln_colsum = colsum.data_ptr()
    args.ln_eps = eps
    return args


def linear_bf16(
    x: Tensor,
    w: Tensor,
    bias: Optional[Tensor] = None,
    *,
    gelu: bool = False,
    scale: float = 1.0,
    resid: Optional[Tensor] = None,
    out_dtype: torch.dtype = torch.bfloat16,
) -> Tensor:
    """``act(x @ w.T * scale + bias) (+ resid)`` for 2-D bf16 ``x`` [M, K] and ``w`` [N, K]."""
    assert x.dim() == 2 and w.dim() == 2 and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    m, k = x.shape
    n = w.shape[0]
    out = torch.empty(m, n, device=x.device, dtype=out_dtype)
    g = make_gemm_args(
        x,
        w,
        a_rows=m,
        a_inner=k,
        a_row_stride=x.stride(0),
        bias=bias,
        gelu=gelu,
        scale=scale,
        resid=resid,
        ld_resid=resid.stride(0) if resid is not None else 0,
        out_f32=out if out_dtype == torch.float32 else None,
        ld_f32=n,
        out_bf16=out if out_dtype == torch.bfloat16 else None,
        ld_bf16=n,
    )
    run_gemm(g)
    return out


# --------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------
class Dropout(NamedTuple):
    """One dropout site of a train-mode step: ``threshold = round(p * 65536)`` (0 = off), the site's 32-bit seed and the
    scale ``1 / (1 - threshold / 65536)`` of the kept values (``aph_common.cuh``: ``drop_hash``)."""

    threshold: int = 0
    seed: int = 0
    scale: float = 1.0

    @classmethod
    def site(cls, probability: float, step_seed: int, site: int) -> "Dropout":
        threshold = int(round(float(probability) * 65536.0))
        if threshold <= 0:
            return cls()
        if threshold >= 65536:
            raise ValueError(f"dropout probability has to be < 1, got {probability}")
        return cls(threshold, mix_seed(step_seed, site), 65536.0 / (65536.0 - threshold))


def _fmix32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def mix_seed(step_seed: int, site: int) -> int:
    """The 32-bit seed of dropout site ``site`` within the step seeded ``step_seed``."""
    return _fmix32((step_seed & 0xFFFFFFFF) ^ _fmix32((site + 1) * 0x9E3779B1))


NO_DROPOUT = Dropout()


def attention(
    q: Tensor, k: Tensor, v: Tensor, ctx: Tensor, frame_lengths: Tensor, n_utt: int, heads: int, seq: int, lse2: Optional[Tensor] = None,
    dropout: Dropout = NO_DROPOUT,
) -> None:  # fmt: skip
    """``q``/``k``/``v`` bf16 ``[n_utt*heads, seq, 64]`` (q pre-scaled) -> ``ctx`` bf16 ``[n_utt*seq, heads*64]``."""
    _require_cuda(q, k, v, ctx, frame_lengths, lse2)
    check(
        lib.aph_attention_bf16_dropout(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), _ptr(lse2), frame_lengths.data_ptr(), n_utt, heads, seq,
            dropout.threshold, dropout.seed, dropout.scale, _stream(),
        ),  # fmt: skip
        "aph_attention_bf16_dropout",
    )


def attention_backward(
    q: Tensor, k: Tensor, v: Tensor, ctx: Tensor, d_ctx: Tensor, lse2: Tensor, delta_scratch: Tensor, dqkv: Tensor,
    frame_lengths: Tensor, n_utt: int, heads: int, seq: int, dropout: Dropout = NO_DROPOUT,
) -> None:  # fmt: skip
    """(dQ | dK | dV) bf16 ``[n_utt*seq, 3*heads*64]`` from the forward's q/k/v/ctx/lse2 and ``d_ctx``."""
    _require_cuda(q, k, v, ctx, d_ctx, lse2, delta_scratch, dqkv, frame_lengths)
    check(
        lib.aph_attention_backward_bf16_dropout(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), d_ctx.data_ptr(), lse2.data_ptr(), delta_scratch.data_ptr(),
            dqkv.data_ptr(), frame_lengths.data_ptr(), n_utt, heads, seq, dropout.threshold, dropout.seed, dropout.scale, _stream(),
        ),  # fmt: skip
        "aph_attention_backward_bf16_dropout",
    )


# --------------------------------------------------------------------------------------
# front end
# --------------------------------------------------------------------------------------
def wave_stats(x: Tensor, lengths: Tensor, stats_scratch: Tensor, mean_rstd: Tensor) -> None:
    _require_cuda(x, lengths, stats_scratch, mean_rstd)
    n, t = x.shape
    check(
        lib.aph_wave_stats(x.data_ptr(), lengths.data_ptr(), n, t, stats_scratch.data_ptr(), mean_rstd.data_ptr(), _stream()),
        "aph_wave_stats",
    )


def zero_mean_unit_var_norm(features: Tensor, lengths: Tensor) -> Tensor:
    """``acoustic_model.py:762-767`` on the GPU (materialised; the fused encoder folds this into conv 0)."""
    _require_cuda(features, lengths)
    features = features.contiguous().float()
    lengths = lengths.to(torch.int64).contiguous()
    n, t = features.shape
    stats = torch.empty(n, 3, device=features.device, dtype=torch.float64)
    mean_rstd = torch.empty(n, 2, device=features.device, dtype=torch.float32)
    wave_stats(features, lengths, stats, mean_rstd)
    out = torch.empty_like(features)
    check(
        lib.aph_wave_norm(features.data_ptr(), lengths.data_ptr(), mean_rstd.data_ptr(), n, t, out.data_ptr(), _stream()),
        "aph_wave_norm",
    )
    return out


def frame_lengths(lengths: Tensor, kernels: Tensor, strides: Tensor, frames32: Optional[Tensor], frames64: Optional[Tensor]) -> None:
    _require_cuda(lengths, kernels, strides, frames32, frames64)
    check(
        lib.aph_frame_lengths(
            lengths.data_ptr(), lengths.numel(), kernels.data_ptr(), strides.data_ptr(), kernels.numel(), _ptr(frames32), _ptr(frames64), _stream()
        ),
        "aph_frame_lengths",
    )


def conv0_ln_gelu(
    x: Tensor,
    lengths: Optional[Tensor],
    mean_rstd: Optional[Tensor],
    w: Tensor,
    bias: Optional[Tensor],
    gamma: Tensor,
    beta: Tensor,
    eps: float,
    out: Tensor,
    skip_padded_frames: bool = True,
) -> None:
    _require_cuda(x, lengths, mean_rstd, w, bias, gamma, beta, out)
    n, t = x.shape
    check(
        lib.aph_conv0_ln_gelu(
            x.data_ptr(), _ptr(lengths), _ptr(mean_rstd), n, t, w.data_ptr(), _ptr(bias), gamma.data_ptr(), beta.data_ptr(), eps, int(skip_padded_frames), out.data_ptr(), _stream()
        ),
        "aph_conv0_ln_gelu",
    )


def conv0_gn_gelu(
    x: Tensor,
    lengths: Optional[Tensor],
    mean_rstd: Optional[Tensor],
    w: Tensor,
    bias: Optional[Tensor],
    gamma: Tensor,
    beta: Tensor,
    eps: float,
    raw_scratch: Tensor,
    stats_scratch: Tensor,
    out: Tensor,
) -> None:
    _require_cuda(x, lengths, mean_rstd, w, bias, gamma, beta, raw_scratch, stats_scratch, out)
    n, t = x.shape
    check(
        lib.aph_conv0_gn_gelu(
            x.data_ptr(),
            _ptr(lengths),
            _ptr(mean_rstd),
            n,
            t,
            w.data_ptr(),
            _ptr(bias),
            gamma.data_ptr(),
            beta.data_ptr(),
            eps,
            raw_scratch.data_ptr(),
            stats_scratch.data_ptr(),
            out.data_ptr(),
            _stream(),
        ),
        "aph_conv0_gn_gelu",
    )


def layernorm_rows(
    x: Tensor,
    rows: int,
    cols: int,
    ld_in: int,
    gamma: Tensor,
    beta: Tensor,
    eps: float,
    *,
    gelu: bool = False,
    out_bf16: Optional[Tensor] = None,
    ld_bf16: int = 0,
    out_f32: Optional[Tensor] = None,
    ld_f32: int = 0,
) -> None:
    _require_cuda(x, gamma, beta, out_bf16, out_f32)
    if cols not in (512, 1024):
        # the row kernels are specialised for the two widths of the XLS-R path; every other width (wav2vec2-base: 768) takes the
        # general kernel of the from-scratch transformer encoder
        if x.dtype != torch.float32 or gelu:
            raise NotImplementedError(f"layernorm_rows over {cols} columns takes fp32 rows and no fused GELU")
        layernorm_any(x, ld_in, rows, cols, gamma, beta, eps, out_f32=out_f32, ld_f32=ld_f32, out_bf16=out_bf16, ld_bf16=ld_bf16)
        return
    check(
        lib.aph_layernorm_rows(
            x.data_ptr(),
            int(x.dtype == torch.float32),
            ld_in,
            rows,
            cols,
            gamma.data_ptr(),
            beta.data_ptr(),
            eps,
            int(gelu),
            _ptr(out_bf16),
            ld_bf16,
            _ptr(out_f32),
            ld_f32,
            _stream(),
        ),
        "aph_layernorm_rows",
    )


# --------------------------------------------------------------------------------------
# heads
# --------------------------------------------------------------------------------------
def compose_embeddings(
    weight: Tensor, tfi: Tensor, category_offsets: Optional[Tensor], rows_out: int, err_flag: Tensor, *, out_bf16: Optional[Tensor] = None, out_f32: Optional[Tensor] = None
) -> None:
    _require_cuda(weight, tfi, category_offsets, err_flag, out_bf16, out_f32)
    v, f = tfi.shape
    check(
        lib.aph_compose_embeddings(
            weight.data_ptr(), weight.shape[0], weight.shape[1], tfi.data_ptr(), _ptr(category_offsets), v, f, rows_out, _ptr(out_bf16), _ptr(out_f32), err_flag.data_ptr(), _stream()
        ),
        "aph_compose_embeddings",
    )


def log_softmax_heads(
    logits: Tensor,
    ld: int,
    rows: int,
    col_lo: int,
    col_span: int,
    col_off: Tensor,
    width: Tensor,
    out_off: Tensor,
    n_heads: int,
    out: Tensor,
    argmax_out: Optional[Tensor],
    maxlp_out: Optional[Tensor],
) -> None:
    _require_cuda(logits, col_off, width, out_off, out, argmax_out, maxlp_out)
    check(
        lib.aph_log_softmax_heads(
            logits.data_ptr(), ld, rows, col_lo, col_span, col_off.data_ptr(), width.data_ptr(), out_off.data_ptr(), n_heads, out.data_ptr(), _ptr(argmax_out), _ptr(maxlp_out), _stream()
        ),
        "aph_log_softmax_heads",
    )


def log_softmax_wide(logits: Tensor, ld: int, rows: int, width: int, out: Tensor, ld_out: int, argmax_out: Optional[Tensor], maxlp_out: Optional[Tensor]) -> None:
    _require_cuda(logits, out, argmax_out, maxlp_out)
    check(
        lib.aph_log_softmax_wide(logits.data_ptr(), ld, rows, width, out.data_ptr(), ld_out, _ptr(argmax_out), _ptr(maxlp_out), _stream()),
        "aph_log_softmax_wide",
    )


def log_softmax(x: Tensor) -> Tensor:
    """``functional.log_softmax(x, -1)`` for an fp32 CUDA tensor of any shape (``acoustic_model.py:1051-1052``)."""
    _require_cuda(x)
    width = x.shape[-1]
    if x.dim() == 3 and x.dtype == torch.float32 and not x.is_contiguous() and x.transpose(0, 1).is_contiguous():
        # the time-first view of a batch-first block (what Allophant.forward returns): normalise the block where it lies and hand
        # back the same view — no transposing copy here, none of the gradient on the way back
        return log_softmax(x.transpose(0, 1)).transpose(0, 1)
    flat = x.float().reshape(-1, width)
    if flat.stride(-1) != 1 or flat.stride(0) != width:
        flat = flat.contiguous()
    out = torch.empty_like(flat)
    log_softmax_wide(flat, width, flat.shape[0], width, out, width, None, None)
    return out.view(x.shape)


def _head_blocks(blocks: Sequence[Tuple]):
    """``(matrix, ld, offset, width[, rows])`` per block -> ctypes array of ``aph_head_block`` (fp32 matrices, the block starts
    ``offset`` elements behind the matrix's first one; ``rows`` = the block's own row count, absent or 0 = the call's)."""
    array = (HeadBlock * len(blocks))()
    for slot, block in zip(array, blocks):
        matrix, ld, offset, width = block[:4]
        if matrix.dtype != torch.float32:
            raise ValueError("head blocks are fp32")
        slot.ptr = matrix.data_ptr() + 4 * offset
        slot.ld = ld
        slot.width = width
        slot.rows = block[4] if len(block) > 4 else 0
    return array


def copy_head_blocks(src: Sequence[Tuple], dst: Sequence[Tuple], rows: int, accumulate: bool = False) -> None:
    """``dst_b (+)= src_b`` for every ``[rows, width]`` block ``(matrix, ld, column, width)`` in one launch."""
    _require_cuda(*[b[0] for b in src], *[b[0] for b in dst])
    if not src:
        return
    check(lib.aph_copy_head_blocks(_head_blocks(src), _head_blocks(dst), len(src), rows, int(accumulate), _stream()), "aph_copy_head_blocks")


def log_softmax_head_blocks(src: Sequence[Tuple[Tensor, int, int, int]], dst: Sequence[Tuple[Tensor, int, int, int]], rows: int) -> None:
    """Row-wise ``log_softmax`` of every ``[rows, width]`` block in one launch."""
    _require_cuda(*[b[0] for b in src], *[b[0] for b in dst])
    if not src:
        return
    check(lib.aph_log_softmax_head_blocks(_head_blocks(src), _head_blocks(dst), len(src), rows, _stream()), "aph_log_softmax_head_blocks")


def log_softmax_many(tensors: Sequence[Tensor]) -> List[Tensor]:
    """``[log_softmax(t, -1) for t in tensors]`` in ONE launch when every tensor is an fp32 ``[A, B, C]`` block that is contiguous
    either as it is or transposed (the time-first views ``Allophant.forward`` returns) and all share ``A * B``; each result has
    the strides of its input.  Anything else goes through :func:`log_softmax` one by one."""
    bases = []
    for t in tensors:
        if t.dim() != 3 or t.dtype != torch.float32 or not t.is_cuda:
            bases = None
            break
        if t.is_contiguous():
            bases.append((t, False))
        elif t.transpose(0, 1).is_contiguous():
            bases.append((t.transpose(0, 1), True))
        else:
            bases = None
            break
    if not bases or len({b.shape[0] * b.shape[1] for b, _ in bases}) != 1:
        return [log_softmax(t) for t in tensors]
    rows = bases[0][0].shape[0] * bases[0][0].shape[1]
    flat = torch.empty(sum(b.numel() for b, _ in bases), device=bases[0][0].device, dtype=torch.float32)
    outputs, src, dst, position = [], [], [], 0
    for base, transposed in bases:
        width = base.shape[2]
        block = flat[position : position + rows * width]
        position += rows * width
        src.append((base, width, 0, width))
        dst.append((block, width, 0, width))
        view = block.view(base.shape)
        outputs.append(view.transpose(0, 1) if transposed else view)
    log_softmax_head_blocks(src, dst, rows)
    return outputs


def dependency_softmax(logits: Tensor, ld: int, rows: int, col_off: Tensor, width: Tensor, dst_col: Tensor, n_deps: int, skip: int, dst: Tensor, ld_dst: int) -> None:
    _require_cuda(logits, col_off, width, dst_col, dst)
    check(
        lib.aph_dependency_softmax(logits.data_ptr(), ld, rows, col_off.data_ptr(), width.data_ptr(), dst_col.data_ptr(), n_deps, skip, dst.data_ptr(), ld_dst, _stream()),
        "aph_dependency_softmax",
    )


def argmax_rows(x: Tensor, ld: int, rows: int, width: int, argmax_out: Tensor, max_out: Tensor) -> None:
    _require_cuda(x, argmax_out, max_out)
    check(lib.aph_argmax_rows(x.data_ptr(), ld, rows, width, argmax_out.data_ptr(), max_out.data_ptr(), _stream()), "aph_argmax_rows")


def ctc_greedy_collapse(
    argmax_in: Tensor, maxlp_in: Tensor, frame_lengths32: Tensor, n_utt: int, seq: int, n_seq: int, blank: int
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Returns (tokens int32 [n_seq, T], timesteps int32 [n_seq, T], counts int32 [n_seq], scores fp32 [n_seq])."""
    _require_cuda(argmax_in, maxlp_in, frame_lengths32)
    dev = argmax_in.device
    tokens = torch.empty(n_seq, seq, device=dev, dtype=torch.int32)
    timesteps = torch.empty(n_seq, seq, device=dev, dtype=torch.int32)
    counts = torch.empty(n_seq, device=dev, dtype=torch.int32)
    scores = torch.empty(n_seq, device=dev, dtype=torch.float32)
    check(
        lib.aph_ctc_greedy_collapse(
            argmax_in.data_ptr(), maxlp_in.data_ptr(), frame_lengths32.data_ptr(), n_utt, seq, n_seq, blank, tokens.data_ptr(), timesteps.data_ptr(), counts.data_ptr(), scores.data_ptr(), _stream()
        ),
        "aph_ctc_greedy_collapse",
    )
    return tokens, timesteps, counts, scores


def ctc_pack_hypotheses(tokens: Tensor, timesteps: Tensor, counts: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(offsets int32 [n_seq+1], packed tokens, packed timesteps): dense copies of the valid prefixes of every row."""
    _require_cuda(tokens, timesteps, counts)
    n_seq, seq = tokens.shape
    dev = tokens.device
    offsets = torch.empty(n_seq + 1, device=dev, dtype=torch.int32)
    packed_tokens = torch.empty(n_seq * seq, device=dev, dtype=torch.int32)
    packed_timesteps = torch.empty(n_seq * seq, device=dev, dtype=torch.int32)
    check(
        lib.aph_ctc_pack_hypotheses(
            tokens.data_ptr(), timesteps.data_ptr(), counts.data_ptr(), n_seq, seq, offsets.data_ptr(), packed_tokens.data_ptr(), packed_timesteps.data_ptr(), _stream()
        ),
        "aph_ctc_pack_hypotheses",
    )
    return offsets, packed_tokens, packed_timesteps


# --------------------------------------------------------------------------------------
# weight packing
# --------------------------------------------------------------------------------------
def cast_bf16(src: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    _require_cuda(src, dst)
    src = src.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    check(lib.aph_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), "aph_cast_bf16")
    return dst


def cast_bf16_2d(src: Tensor, ld_src: int, dst: Tensor, ld_dst: int, rows: int, cols: int) -> None:
    _require_cuda(src, dst)
    check(lib.aph_cast_bf16_2d(src.data_ptr(), ld_src, dst.data_ptr(), ld_dst, rows, cols, _stream()), "aph_cast_bf16_2d")


def dropout_2d(
    x: Tensor, ld_x: int, rows: int, cols: int, dropout: Dropout, out_f32: Optional[Tensor] = None, ld_f32: int = 0,
    out_bf16: Optional[Tensor] = None, ld_bf16: int = 0, row_mask: Optional[Tensor] = None, row_fill: Optional[Tensor] = None,
) -> None:  # fmt: skip
    """``out = x * keep * scale`` (fp32 ``x``; ``out_f32`` may be ``x``), rows flagged in ``row_mask`` replaced by ``row_fill``."""
    _require_cuda(x, out_f32, out_bf16, row_mask, row_fill)
    check(
        lib.aph_dropout_2d(
            x.data_ptr(), ld_x, rows, cols, dropout.threshold, dropout.seed, dropout.scale, _ptr(row_mask), _ptr(row_fill),
            _ptr(out_f32), ld_f32, _ptr(out_bf16), ld_bf16, _stream(),
        ),  # fmt: skip
        "aph_dropout_2d",
    )


def dropout_bf16_2d(x: Tensor, ld: int, rows: int, cols: int, dropout: Dropout) -> None:
    _require_cuda(x)
    check(lib.aph_dropout_bf16_2d(x.data_ptr(), ld, rows, cols, dropout.threshold, dropout.seed, dropout.scale, _stream()), "aph_dropout_bf16_2d")


def spec_augment_mask(frames32: Tensor, seq: int, mask_prob: float, mask_length: int, min_masks: int, seed: int, mask: Tensor) -> None:
    """uint8 ``mask`` ``[n_utt, seq]`` <- SpecAugment time mask (HF ``_compute_mask_indices``)."""
    _require_cuda(frames32, mask)
    if mask_length > seq:
        raise ValueError(f"`mask_length` has to be smaller than `sequence_length`, but got `mask_length`: {mask_length} and `sequence_length`: {seq}`")
    check(
        lib.aph_spec_augment_mask(frames32.data_ptr(), frames32.numel(), seq, mask_prob, mask_length, min_masks, seed & 0xFFFFFFFF, mask.data_ptr(), _stream()),
        "aph_spec_augment_mask",
    )


def mask_columns(x: Tensor, ld: int, n_utt: int, seq: int, cols: int, col_mask: Tensor, x_bf16: Optional[Tensor] = None, ld_bf16: int = 0) -> None:
    """``x[n, t, c] = 0`` where ``col_mask[n, c]`` (SpecAugment along the feature axis); fp32 in place + optional bf16 copy."""
    _require_cuda(x, col_mask, x_bf16)
    check(lib.aph_mask_columns(x.data_ptr(), ld, n_utt, seq, cols, col_mask.data_ptr(), _ptr(x_bf16), ld_bf16, _stream()), "aph_mask_columns")


def masked_rows_backward(d: Tensor, ld: int, rows: int, cols: int, row_mask: Tensor, d_fill: Tensor) -> None:
    _require_cuda(d, row_mask, d_fill)
    check(lib.aph_masked_rows_backward(d.data_ptr(), ld, rows, cols, row_mask.data_ptr(), d_fill.data_ptr(), _stream()), "aph_masked_rows_backward")


def conv0_raw(audio: Tensor, lengths: Tensor, mean_rstd: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], out: Tensor) -> None:
    """Pre-LayerNorm output of the first wav2vec2 convolution, bf16 ``[N, L0, 512]`` (feature-extractor training)."""
    _require_cuda(audio, lengths, mean_rstd, weight, bias, out)
    check(lib.aph_conv0_raw_bf16(audio.data_ptr(), lengths.data_ptr(), _ptr(mean_rstd), audio.shape[0], audio.shape[1], weight.data_ptr(), _ptr(bias), out.data_ptr(), _stream()), "aph_conv0_raw_bf16")


def ln_gelu_backward_512(x: Tensor, d_out: Tensor, ld_d: int, rows: int, gamma: Tensor, beta: Tensor, eps: float, dx: Tensor, dgamma: Tensor, dbeta: Tensor, dbias: Optional[Tensor]) -> None:
    _require_cuda(x, d_out, gamma, beta, dx, dgamma, dbeta, dbias)
    check(lib.aph_ln_gelu_backward_512(x.data_ptr(), d_out.data_ptr(), ld_d, rows, gamma.data_ptr(), beta.data_ptr(), eps, dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _ptr(dbias), _stream()), "aph_ln_gelu_backward_512")


def conv0_weight_backward(dy: Tensor, audio: Tensor, lengths: Tensor, mean_rstd: Optional[Tensor], dw: Tensor) -> None:
    _require_cuda(dy, audio, lengths, mean_rstd, dw)
    check(lib.aph_conv0_weight_backward(dy.data_ptr(), audio.data_ptr(), lengths.data_ptr(), _ptr(mean_rstd), audio.shape[0], audio.shape[1], dw.data_ptr(), _stream()), "aph_conv0_weight_backward")


def layernorm_any(
    x: Tensor, ld_x: int, rows: int, cols: int, gamma: Optional[Tensor], beta: Optional[Tensor], eps: float,
    out_f32: Optional[Tensor] = None, ld_f32: int = 0, out_bf16: Optional[Tensor] = None, ld_bf16: int = 0,
) -> None:  # fmt: skip
    """``nn.LayerNorm`` over the last axis of fp32 ``x`` (any width; ``gamma``/``beta`` ``None`` = no affine parameters)."""
    _require_cuda(x, gamma, beta, out_f32, out_bf16)
    check(
        lib.aph_layernorm_any(x.data_ptr(), ld_x, rows, cols, _ptr(gamma), _ptr(beta), eps, _ptr(out_f32), ld_f32, _ptr(out_bf16), ld_bf16, _stream()),
        "aph_layernorm_any",
    )


def layernorm_any_backward(
    x: Tensor, ld_x: int, dy: Tensor, ld_dy: int, rows: int, cols: int, gamma: Optional[Tensor], eps: float,
    resid: Optional[Tensor], ld_resid: int, dx: Optional[Tensor], ld_dx: int, dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None,
) -> None:  # fmt: skip
    """Backward of ``layernorm_any``: ``dx = LN'(dy) (+ resid)``; ``dgamma`` / ``dbeta`` are accumulated into."""
    _require_cuda(x, dy, gamma, resid, dx, dgamma, dbeta)
    check(
        lib.aph_layernorm_any_backward(
            x.data_ptr(), ld_x, dy.data_ptr(), ld_dy, rows, cols, _ptr(gamma), eps, _ptr(resid), ld_resid, _ptr(dx), ld_dx, _ptr(dgamma), _ptr(dbeta), _stream()
        ),
        "aph_layernorm_any_backward",
    )


def glu_backward_bf16(y: Tensor, ld_y: int, d_out: Tensor, ld_d: int, rows: int, out_channels: int, dy: Tensor, ld_dy: int) -> None:
    _require_cuda(y, d_out, dy)
    check(lib.aph_glu_backward_bf16(y.data_ptr(), ld_y, d_out.data_ptr(), ld_d, rows, out_channels, dy.data_ptr(), ld_dy, _stream()), "aph_glu_backward_bf16")


def conv_input_backward(
    d_cols: Tensor, lengths32: Tensor, n_utt: int, length: int, channels: int, out_len: int, kernel: int, stride: int, left: int, right: int,
    reflect: bool, d_x: Tensor, ld_dx: int,
) -> None:  # fmt: skip
    _require_cuda(d_cols, lengths32, d_x)
    check(
        lib.aph_conv_input_backward(d_cols.data_ptr(), lengths32.data_ptr(), n_utt, length, channels, out_len, kernel, stride, left, right, int(reflect), d_x.data_ptr(), ld_dx, _stream()),
        "aph_conv_input_backward",
    )


def activation_backward(d: Tensor, ld_d: int, y: Tensor, ld_y: int, rows: int, cols: int, kind: int, out_bf16: Optional[Tensor] = None, ld_bf16: int = 0) -> None:
    """``d *= act'`` decided from the activation output ``y`` (kind 2 ReLU, 3 LeakyReLU(0.01)); optional bf16 copy."""
    _require_cuda(d, y, out_bf16)
    check(lib.aph_activation_backward(d.data_ptr(), ld_d, y.data_ptr(), ld_y, rows, cols, kind, _ptr(out_bf16), ld_bf16, _stream()), "aph_activation_backward")


def attention_small(qkv: Tensor, ld: int, ctx: Tensor, ld_ctx: int, lengths: Tensor, n_utt: int, heads: int, seq: int, head_dim: int) -> None:
    """Self-attention over time for the time layer of a classifier head (any head_dim): ``qkv`` fp32 ``[n_utt*seq, ld]`` holds
    q | k | v, ``ctx`` bf16 ``[n_utt*seq, ld_ctx]`` receives the per-head outputs; keys at or beyond ``lengths[n]`` are masked."""
    _require_cuda(qkv, ctx, lengths)
    check(lib.aph_attention_small(qkv.data_ptr(), ld, ctx.data_ptr(), ld_ctx, lengths.data_ptr(), n_utt, heads, seq, head_dim, _stream()), "aph_attention_small")


def add_sinusoidal(x: Tensor, ld: int, n_utt: int, seq: int, cols: int, bases: Tensor) -> None:
    _require_cuda(x, bases)
    check(lib.aph_add_sinusoidal(x.data_ptr(), ld, n_utt, seq, cols, bases.data_ptr(), _stream()), "aph_add_sinusoidal")


def transpose_nfl(features: Tensor, out: Tensor, ld_out: int) -> None:
    """fp32 ``[N, F, L]`` -> channels-last ``[N, L, ld_out]``."""
    _require_cuda(features, out)
    n, f, length = features.shape
    check(lib.aph_transpose_nfl(features.data_ptr(), n, f, length, out.data_ptr(), ld_out, _stream()), "aph_transpose_nfl")


def reflect_pad_bf16(x: Tensor, ld_x: int, lengths32: Tensor, n_utt: int, length: int, channels: int, left: int, right: int, reflect: bool, out: Tensor) -> None:
    _require_cuda(x, lengths32, out)
    check(
        lib.aph_reflect_pad_bf16(x.data_ptr(), ld_x, lengths32.data_ptr(), n_utt, length, channels, left, right, int(reflect), out.data_ptr(), _stream()),
        "aph_reflect_pad_bf16",
    )


def glu_rows(y: Tensor, ld_y: int, rows: int, out_channels: int, out: Tensor, ld_out: int) -> None:
    _require_cuda(y, out)
    check(lib.aph_glu_rows(y.data_ptr(), ld_y, rows, out_channels, out.data_ptr(), ld_out, _stream()), "aph_glu_rows")


def pack_conv_weight(weight: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    """Conv1d weight [O, C, k] fp32 -> bf16 [O, k*C]."""
    _require_cuda(weight, dst)
    w = weight.detach().float().contiguous()
    o, c, k = w.shape
    if dst is None:
        dst = torch.empty(o, k * c, device=w.device, dtype=torch.bfloat16)
    check(lib.aph_pack_conv_weight(w.data_ptr(), dst.data_ptr(), o, c, k, _stream()), "aph_pack_conv_weight")
    return dst


def pack_posconv_weight(weight_g: Tensor, weight_v: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    """weight-normed grouped conv weight -> bf16 [O, k*Cg] (tap-major)."""
    _require_cuda(weight_g, weight_v, dst)
    g = weight_g.detach().float().contiguous()
    v = weight_v.detach().float().contiguous()
    o, cg, k = v.shape
    if dst is None:
        dst = torch.empty(o, k * cg, device=v.device, dtype=torch.bfloat16)
    scratch = torch.empty(k, device=v.device, dtype=torch.float32)
    check(lib.aph_pack_posconv_weight(g.data_ptr(), v.data_ptr(), dst.data_ptr(), scratch.data_ptr(), o, cg, k, _stream()), "aph_pack_posconv_weight")
    return dst


# --------------------------------------------------------------------------------------
# training-side kernels of the classifier heads
# --------------------------------------------------------------------------------------
def allophone_forward(logits: Tensor, matrices: Tensor, csr_offsets: Tensor, csr_phones: Tensor, language_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """``logits`` fp32 [N, T, P+1] (any N/T strides) -> (mapped [N, T, Q+1] fp32, argmax int32)."""
    _require_cuda(logits, matrices, csr_offsets, csr_phones, language_ids)
    n, t, p1 = logits.shape
    q1 = matrices.shape[2]
    out = torch.empty(n, t, q1, device=logits.device, dtype=torch.float32)
    arg = torch.empty(n, t, q1, device=logits.device, dtype=torch.int32)
    check(
        lib.aph_allophone_forward(
            logits.data_ptr(), logits.stride(0), logits.stride(1), n, t, p1, q1, matrices.data_ptr(), csr_offsets.data_ptr(), csr_phones.data_ptr(), language_ids.data_ptr(), out.data_ptr(), arg.data_ptr(), _stream()
        ),
        "aph_allophone_forward",
    )
    return out, arg


def allophone_backward(grad_out: Tensor, arg: Tensor, logits: Tensor, matrices: Tensor, language_ids: Tensor, need_logits_grad: bool, need_matrix_grad: bool):
    _require_cuda(grad_out, arg, logits, matrices, language_ids)
    n, t, p1 = logits.shape
    q1 = matrices.shape[2]
    grad_logits = torch.zeros(n, t, p1, device=logits.device, dtype=torch.float32) if need_logits_grad else None
    grad_matrices = torch.zeros_like(matrices, dtype=torch.float32) if need_matrix_grad else None
    check(
        lib.aph_allophone_backward(
            grad_out.data_ptr(), arg.data_ptr(), logits.data_ptr(), logits.stride(0), logits.stride(1), n, t, p1, q1, matrices.data_ptr(), language_ids.data_ptr(), _ptr(grad_logits), _ptr(grad_matrices), _stream()
        ),
        "aph_allophone_backward",
    )
    return grad_logits, grad_matrices


def transpose_cast_bf16(x: Tensor, rows: int, cols: int, ld_in: int, rows_padded: int) -> Tensor:
    """[rows, cols] (fp32/bf16, row stride ``ld_in``) -> bf16 [cols, rows_padded], zero padded."""
    _require_cuda(x)
    out = torch.empty(cols, rows_padded, device=x.device, dtype=torch.bfloat16)
    check(lib.aph_transpose_cast_bf16(x.data_ptr(), int(x.dtype == torch.float32), ld_in, rows, cols, out.data_ptr(), rows_padded, rows_padded, _stream()), "aph_transpose_cast_bf16")
    return out


def colsum_f32(x: Tensor, rows: int, cols: int, ld: int, out: Optional[Tensor] = None) -> Tensor:
    _require_cuda(x, out)
    if out is None:
        out = torch.empty(cols, device=x.device, dtype=torch.float32)
    check(lib.aph_colsum_f32(x.data_ptr(), ld, rows, cols, out.data_ptr(), _stream()), "aph_colsum_f32")
    return out


def colsum_bf16(x: Tensor, rows: int, cols: int, ld: int, out: Optional[Tensor] = None) -> Tensor:
    _require_cuda(x, out)
    if out is None:
        out = torch.empty(cols, device=x.device, dtype=torch.float32)
    check(lib.aph_colsum_bf16(x.data_ptr(), ld, rows, cols, out.data_ptr(), _stream()), "aph_colsum_bf16")
    return out


def layernorm_backward(
    x: Tensor, ld_x: int, dy: Tensor, ld_dy: int, rows: int, cols: int, gamma: Tensor, eps: float,
    dx_resid: Optional[Tensor], ld_resid: int, dx: Tensor, ld_dx: int, dgamma: Optional[Tensor], dbeta: Optional[Tensor],
) -> None:  # fmt: skip
    _require_cuda(x, dy, gamma, dx_resid, dx, dgamma, dbeta)
    if cols not in (512, 1024):
        if x.dtype != torch.float32 or dy.dtype != torch.float32:
            raise NotImplementedError(f"layernorm_backward over {cols} columns takes fp32 rows")
        if dgamma is not None:  # the general kernel accumulates into them
            dgamma.zero_()
            dbeta.zero_()
        layernorm_any_backward(x, ld_x, dy, ld_dy, rows, cols, gamma, eps, dx_resid, ld_resid, dx, ld_dx, dgamma, dbeta)
        return
    check(
        lib.aph_layernorm_backward(
            x.data_ptr(), int(x.dtype == torch.float32), ld_x, dy.data_ptr(), int(dy.dtype == torch.float32), ld_dy, rows, cols,
            gamma.data_ptr(), eps, _ptr(dx_resid), ld_resid, dx.data_ptr(), ld_dx, _ptr(dgamma), _ptr(dbeta), _stream(),
        ),  # fmt: skip
        "aph_layernorm_backward",
    )


def mask_rows(x: Tensor, ld: int, rows: int, cols: int, frame_lengths32: Tensor, period: int) -> None:
    _require_cuda(x, frame_lengths32)
    check(lib.aph_mask_rows_f32(x.data_ptr(), ld, rows, cols, frame_lengths32.data_ptr(), period, _stream()), "aph_mask_rows_f32")


def add_2d(dst: Tensor, ld_dst: int, src: Tensor, ld_src: int, rows: int, cols: int) -> None:
    _require_cuda(dst, src)
    check(lib.aph_add_f32_2d(dst.data_ptr(), ld_dst, src.data_ptr(), ld_src, rows, cols, _stream()), "aph_add_f32_2d")


def gelu_backward_bf16(dy: Tensor, ld_dy: int, pre: Tensor, ld_pre: int, rows: int, cols: int, out: Tensor, ld_out: int) -> None:
    _require_cuda(dy, pre, out)
    check(lib.aph_gelu_backward_bf16(dy.data_ptr(), ld_dy, pre.data_ptr(), ld_pre, rows, cols, out.data_ptr(), ld_out, _stream()), "aph_gelu_backward_bf16")


def pack_posconv_weight_dgrad(weight_g: Tensor, weight_v: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    """B operand of the positional conv's data-gradient GEMM (taps flipped, in/out channels swapped)."""
    _require_cuda(weight_g, weight_v, dst)
    g = weight_g.detach().float().contiguous()
    v = weight_v.detach().float().contiguous()
    o, cg, k = v.shape
    if dst is None:
        dst = torch.empty(o, k * cg, device=v.device, dtype=torch.bfloat16)
    scratch = torch.empty(2 * k, device=v.device, dtype=torch.float32)
    check(lib.aph_pack_posconv_weight_dgrad(g.data_ptr(), v.data_ptr(), dst.data_ptr(), scratch.data_ptr(), o, cg, k, _stream()), "aph_pack_posconv_weight_dgrad")
    return dst


def posconv_weight_backward(
    raw: Tensor, weight_g: Tensor, weight_v: Tensor, grad_g: Optional[Tensor] = None, grad_v: Optional[Tensor] = None, block_width: int = 256
) -> Tuple[Tensor, Tensor]:
    """``raw`` fp32 [k, O, block_width] (256: from the DIAG_TAPS GEMM; O: one full weight-gradient GEMM per tap) -> (grad of
    original0 [1,1,k], grad of original1 [O,Cg,k])."""
    _require_cuda(raw, weight_g, weight_v, grad_g, grad_v)
    g = weight_g.detach().float().contiguous()
    v = weight_v.detach().float().contiguous()
    o, cg, k = v.shape
    if grad_g is None:
        grad_g = torch.empty(g.shape, device=v.device, dtype=torch.float32)
    if grad_v is None:
        grad_v = torch.empty(v.shape, device=v.device, dtype=torch.float32)
    scratch = torch.empty(3 * k, device=v.device, dtype=torch.float32)
    check(
        lib.aph_posconv_weight_backward_blocks(raw.data_ptr(), g.data_ptr(), v.data_ptr(), scratch.data_ptr(), o, cg, k, block_width, grad_g.data_ptr(), grad_v.data_ptr(), _stream()),
        "aph_posconv_weight_backward_blocks",
    )
    return grad_g, grad_v


def embedding_bag_backward(grad_rows: Tensor, ld: int, tfi: Tensor, category_offsets: Optional[Tensor], grad_weight: Tensor) -> None:
    _require_cuda(grad_rows, tfi, category_offsets, grad_weight)
    v, f = tfi.shape
    check(lib.aph_embedding_bag_backward(grad_rows.data_ptr(), ld, v, f, grad_weight.shape[1], tfi.data_ptr(), _ptr(category_offsets), grad_weight.data_ptr(), _stream()), "aph_embedding_bag_backward")


def softmax_backward_cols(grad_x: Tensor, ld_gx: int, x: Tensor, ld_x: int, rows: int, x_col: Tensor, width: Tensor, dst_col: Tensor, n_deps: int, skip: int, grad_logits: Tensor, ld_gl: int) -> None:
    _require_cuda(grad_x, x, x_col, width, dst_col, grad_logits)
    check(
        lib.aph_softmax_backward_cols(grad_x.data_ptr(), ld_gx, x.data_ptr(), ld_x, rows, x_col.data_ptr(), width.data_ptr(), dst_col.data_ptr(), n_deps, skip, grad_logits.data_ptr(), ld_gl, _stream()),
        "aph_softmax_backward_cols",
    )


# --------------------------------------------------------------------------------------
# CTC
# --------------------------------------------------------------------------------------
class CtcProblem:
    """Device + host copies of the per-head descriptors of one multi-head CTC evaluation."""

    def __init__(
        self,
        log_probs: Sequence[Tensor],
        labels: Sequence[Tensor],
        label_lengths: Sequence[Tensor],
        input_lengths: Tensor,
        batch_first: bool,
        need_grad: bool,
    ) -> None:
        _require_cuda(input_lengths, *log_probs, *labels, *label_lengths)
        self.n_heads = len(log_probs)
        first = log_probs[0]
        self.n_utt = first.shape[0] if batch_first else first.shape[1]
        self.seq = first.shape[1] if batch_first else first.shape[0]
        dev = first.device
        self.input_lengths = input_lengths.to(device=dev, dtype=torch.int64).contiguous()
        self.max_label_len = max(int(l.shape[1]) if l.dim() == 2 else 0 for l in labels)
        s_pad = int(lib.aph_ctc_states_pad(self.max_label_len))
        if s_pad < 0:
            raise NotImplementedError(f"CTC label sequences longer than 8000 are not supported (got {self.max_label_len})")
        self.s_pad = s_pad
        self.keep: List[Tensor] = []
        self.grads: List[Optional[Tensor]] = []
        heads = (CtcHead * self.n_heads)()
        alpha_total = 0
        for h, (lp, lab, lab_len) in enumerate(zip(log_probs, labels, label_lengths)):
            if lp.dtype != torch.float32 or lp.stride(-1) != 1:
                raise ValueError("CTC log-probabilities must be fp32 with a contiguous class axis")
            lab = lab.to(device=dev, dtype=torch.int64)
            if lab.dim() != 2:
                raise ValueError("CTC labels must be a padded [N, S_max] matrix")
            if lab.shape[1] == 0:
                lab = torch.zeros(self.n_utt, 1, device=dev, dtype=torch.int64)
            lab = lab.contiguous()
            lab_len = lab_len.to(device=dev, dtype=torch.int64).contiguous()
            grad = torch.empty_like(lp, memory_format=torch.preserve_format) if need_grad else None
            if grad is not None and grad.stride() != lp.stride():
                grad = torch.empty_strided(lp.shape, lp.stride(), device=dev, dtype=lp.dtype)
            self.keep += [lp, lab, lab_len]
            self.grads.append(grad)
            hd = heads[h]
            hd.log_probs = lp.data_ptr()
            hd.grad = _ptr(grad)
            hd.stride_n, hd.stride_t = (lp.stride(0), lp.stride(1)) if batch_first else (lp.stride(1), lp.stride(0))
            hd.n_classes = lp.shape[2]
            hd.s_pad = s_pad
            hd.labels = lab.data_ptr()
            hd.label_stride = lab.stride(0)
            hd.label_lengths = lab_len.data_ptr()
            hd.alpha_offset = alpha_total
            alpha_total += self.n_utt * self.seq * s_pad
        self.heads_host = heads  # passed by value in the kernel parameters: no device copy, no synchronous transfer
        self.alpha = torch.empty(alpha_total, device=dev, dtype=torch.float32) if need_grad else None
        self.nll = torch.empty(self.n_heads, self.n_utt, device=dev, dtype=torch.float32)
        self.loss = torch.empty(self.n_heads, device=dev, dtype=torch.float32)

    def forward(self) -> Tensor:
        check(
            lib.aph_ctc_forward(
                self.heads_host, self.n_heads, self.n_utt, self.seq, self.max_label_len, self.input_lengths.data_ptr(), _ptr(self.alpha), self.nll.data_ptr(), self.loss.data_ptr(), _stream()
            ),
            "aph_ctc_forward",
        )
        return self.loss

    def backward(self, grad_scale: Tensor) -> List[Optional[Tensor]]:
        if self.alpha is None:
            raise RuntimeError("CtcProblem was built without need_grad")
        grad_scale = grad_scale.to(dtype=torch.float32).contiguous()
        check(
            lib.aph_ctc_backward(
                self.heads_host, self.n_heads, self.n_utt, self.seq, self.max_label_len, self.input_lengths.data_ptr(), self.alpha.data_ptr(), self.nll.data_ptr(), grad_scale.data_ptr(), _stream()
            ),
            "aph_ctc_backward",
        )
        return self.grads
