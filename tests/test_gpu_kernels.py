"""-m gpu: every C-ABI kernel against the CPU oracle / torch's own CPU ops on seeded inputs.

Tolerances: integer / index results bit-exact; fp32 kernels 1e-3 relative (north_star) — in practice
they sit at 1e-6; kernels with bf16 operands 2e-2 of the tensor's range.
"""
import math

import pytest
import torch
from torch.nn import functional as F

from oracle import restatement

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from allophant_b200 import ops

    return ops


def range_err(value, reference):
    reference = reference.double().cpu()
    return float((value.double().cpu() - reference).abs().max() / reference.abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1000, 1024, 1024), (333, 784, 1024), (77, 32, 640), (4096, 4096, 1024), (150, 64, 40), (300, 256, 200)])
def test_gemm_plain(m, n, k):
    ops = _ops()
    torch.manual_seed(m + n + k)
    a = (torch.randn(m, k, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(n, device=DEV)
    out = ops.linear_bf16(a, w, bias, out_dtype=torch.float32)
    reference = a.float() @ w.float().T + bias
    assert range_err(out, reference) < 1e-4  # same bf16 inputs, fp32 accumulate: only summation order differs
    out16 = ops.linear_bf16(a, w, bias, gelu=True)
    assert range_err(out16, F.gelu(reference)) < 1e-2


def test_gemm_residual_mask_and_scale():
    ops = _ops()
    torch.manual_seed(1)
    m, n, k, period = 998, 1024, 512, 499
    a = (torch.randn(m, k, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(n, device=DEV)
    resid = torch.randn(m, n, device=DEV)
    lengths = torch.tensor([300, 1], device=DEV, dtype=torch.int32)
    out = resid.clone()
    args = ops.make_gemm_args(
        a, w, a_rows=m, a_inner=k, a_row_stride=k, bias=bias, scale=0.5, resid=out, ld_resid=n, out_f32=out, ld_f32=n, lengths=lengths, len_period=period
    )
    ops.run_gemm(args)
    reference = (a.float() @ w.float().T) * 0.5 + bias + resid
    rows = torch.arange(m, device=DEV)
    reference[(rows % period) >= lengths[rows // period]] = 0
    assert range_err(out, reference) < 1e-4


def test_gelu_matches_erf():
    """The GELU of every epilogue / row kernel (aph_common.cuh:gelu_erf*: erfc as 2^R(-|x|) with a degree-5 fit of log2 erfc)
    against the exact erf form in fp64, through a GEMM with an identity weight (its product reproduces the bf16 input exactly)."""
    ops = _ops()
    values = torch.linspace(-9.0, 9.0, 64 * 4096, device=DEV).bfloat16()
    x = values.view(4096, 64)
    identity = torch.eye(64, device=DEV).bfloat16()
    out = torch.empty(4096, 64, device=DEV)
    ops.run_gemm(ops.make_gemm_args(x, identity, a_rows=4096, a_inner=64, a_row_stride=64, gelu=True, out_f32=out, ld_f32=64))
    exact = F.gelu(x.double())
    error = (out.double() - exact).abs()
    assert float(error.max()) < 2.5e-6, float(error.max())
    # the negative tail keeps its relative accuracy well inside bf16 resolution (3.9e-3)
    tail = exact.abs() > 1e-3
    assert float((error[tail] / exact[tail].abs()).max()) < 1.5e-3


@pytest.mark.parametrize(
    "m,n,k,form",
    [
        (1024, 1024, 1024, "residual"),   # 16 tiles on 74 cluster slots: every tile becomes two half-width items
        (5120, 1024, 1024, "residual"),   # 80 tiles: one full wave + 6 tail tiles (the out-proj of a training step)
        (4904, 1024, 4096, "residual"),   # ragged last row tile, K = 4096 (FFN2)
        (16384, 1024, 1024, "residual"),  # 256 tiles: 3 full waves + 34 tail tiles (out-proj of the 32 x 10 s batch)
        (5120, 4096, 1024, "gelu_bf16"),  # 320 tiles: 4 waves + 24 tail tiles, bf16 output through 64-column stores
        (4904, 1024, 3072, "dgrad"),      # MN-major B, plain fp32 store
        (2048, 512, 520, "plain"),        # odd number of k-blocks (9)
    ],
)
def test_gemm_tail_split_matches_the_unsplit_kernel(m, n, k, form):
    """The tiles of a partly filled last wave are computed as two half-width work items on two clusters: every output
    element accumulates the same products in the same order, so the result equals the unsplit kernel's bit for bit."""
    ops = _ops()
    torch.manual_seed(m + k)
    a = (torch.randn(m, k, device=DEV) * 0.5).bfloat16()
    bias = torch.randn(n, device=DEV)
    resid = torch.randn(m, n, device=DEV)

    def run():
        if form == "dgrad":
            w = (torch.randn(k, n, device=DEV) * 0.05).bfloat16()  # stored [out = k][in = n], as the forward pass keeps it
            out = torch.empty(m, n, device=DEV)
            ops.run_gemm(ops.make_dgrad_args(a, w, rows=m, ld_dy=k, k=k, n=n, ld_w=n, out_f32=out, ld_f32=n))
            return out, a.float() @ w.float()
        w = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
        product = a.float() @ w.float().T
        if form == "residual":
            out = resid.clone()
            ops.run_gemm(ops.make_gemm_args(a, w, a_rows=m, a_inner=k, a_row_stride=k, bias=bias, resid=out, ld_resid=n, out_f32=out, ld_f32=n))
            return out, product + bias + resid
        if form == "gelu_bf16":
            out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
            ops.run_gemm(ops.make_gemm_args(a, w, a_rows=m, a_inner=k, a_row_stride=k, bias=bias, gelu=True, out_bf16=out, ld_bf16=n))
            return out.float(), torch.nn.functional.gelu(product + bias)
        out = torch.empty(m, n, device=DEV)
        ops.run_gemm(ops.make_gemm_args(a, w, a_rows=m, a_inner=k, a_row_stride=k, out_f32=out, ld_f32=n))
        return out, product

    before = ops.set_gemm_tail_split(True)
    try:
        torch.manual_seed(7)
        split, reference = run()
        torch.manual_seed(7)
        again, _ = run()
        ops.set_gemm_tail_split(False)
        torch.manual_seed(7)
        whole, _ = run()
    finally:
        ops.set_gemm_tail_split(before)
    assert torch.equal(split, again)
    assert torch.equal(split, whole)
    assert range_err(split, reference) < (1e-2 if form == "gelu_bf16" else 1e-4)


@pytest.mark.parametrize("m,period,lengths", [(998, 499, [300, 499]), (130, 130, [130]), (1537, 1537, [1500])])
def test_gemm_folded_layernorm_pair(m, period, lengths):
    """LayerNorm folded into the GEMMs around it (HF:730-756, inference launch list): the producer (out-proj form: bias +
    residual, fp32 in place) leaves a bf16 copy and per-row (sum, sum of squares); the consumers (FFN1 form: GELU, bf16 out;
    q/k/v form: scatter epilogue) read the copy and a gamma-folded weight and apply mean / rstd themselves."""
    ops = _ops()
    torch.manual_seed(m)
    h, ff, heads = 1024, 4096, 16
    ctx = (torch.randn(m, h, device=DEV) * 0.5).bfloat16()
    wo = (torch.randn(h, h, device=DEV) * 0.03).bfloat16()
    bo = torch.randn(h, device=DEV)
    resid = torch.randn(m, h, device=DEV) * 2 + 0.7  # a row mean that is not zero
    resid[:, 5] += 40.0  # and an outlier channel, as real wav2vec2 residual streams have
    lens = torch.tensor(lengths, device=DEV, dtype=torch.int32)
    hidden = resid.clone()
    copy16 = torch.zeros(m, h, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(m, 2 * (h // 256), 2, device=DEV)
    producer = ops.make_gemm_args(ctx, wo, a_rows=m, a_inner=h, a_row_stride=h, bias=bo, resid=hidden, ld_resid=h, out_f32=hidden, ld_f32=h,
                                  out_bf16=copy16, ld_bf16=h, lengths=lens, len_period=period)  # fmt: skip
    ops.run_gemm(ops.with_row_stats(producer, stats))
    expected_hidden = ctx.float() @ wo.float().T + bo + resid
    rows = torch.arange(m, device=DEV)
    expected_hidden[(rows % period) >= lens[rows // period]] = 0
    assert range_err(hidden, expected_hidden) < 1e-4
    assert torch.equal(copy16, hidden.bfloat16())
    sums = stats.sum(1)
    assert range_err(sums[:, 0], hidden.sum(1)) < 1e-5 and range_err(sums[:, 1], (hidden * hidden).sum(1)) < 1e-5

    gamma, beta = torch.rand(h, device=DEV) + 0.5, torch.randn(h, device=DEV) * 0.3
    normalised = F.layer_norm(hidden, (h,), gamma, beta, 1e-5)
    # FFN1 form
    w1, b1 = torch.randn(ff, h, device=DEV) * 0.03, torch.randn(ff, device=DEV)
    w1_folded, colsum, bias_folded = torch.empty(ff, h, device=DEV, dtype=torch.bfloat16), torch.empty(ff, device=DEV), torch.empty(ff, device=DEV)
    ops.fold_layernorm_linear(w1, b1, gamma, beta, w1_folded, colsum, bias_folded)
    assert torch.equal(w1_folded, (w1 * gamma).bfloat16())
    assert range_err(colsum, w1_folded.float().sum(1)) < 1e-5 and range_err(bias_folded, b1 + w1 @ beta) < 1e-5
    act = torch.zeros(m, ff, device=DEV, dtype=torch.bfloat16)
    consumer = ops.make_gemm_args(copy16, w1_folded, a_rows=m, a_inner=h, a_row_stride=h, bias=bias_folded, gelu=True, out_bf16=act, ld_bf16=ff)
    ops.run_gemm(ops.with_layernorm(consumer, stats, colsum, h, 1e-5))
    reference = F.gelu(normalised @ w1.T + b1)
    unfused = F.gelu(normalised.bfloat16().float() @ w1.bfloat16().float().T + b1)  # what the LayerNorm kernel + plain GEMM compute
    assert range_err(act, reference) < 1e-2
    assert range_err(act, reference) < 2.0 * range_err(unfused.bfloat16(), reference) + 1e-3  # as close to fp32 as the unfused path
    # q/k/v form
    wqkv, bqkv = torch.randn(3 * h, h, device=DEV) * 0.03, torch.randn(3 * h, device=DEV)
    wq_folded, sq, bq = torch.empty(3 * h, h, device=DEV, dtype=torch.bfloat16), torch.empty(3 * h, device=DEV), torch.empty(3 * h, device=DEV)
    ops.fold_layernorm_linear(wqkv, bqkv, gamma, beta, wq_folded, sq, bq)
    n_utt = m // period
    q, k, v = (torch.zeros(n_utt * heads, period, 64, device=DEV, dtype=torch.bfloat16) for _ in range(3))
    args = ops.make_qkv_args(copy16, wq_folded, bq, q, k, v, rows=m, seq=period, heads=heads)
    ops.run_gemm(ops.with_layernorm(args, stats, sq, h, 1e-5))
    projected = (normalised @ wqkv.T + bqkv).view(n_utt, period, 3, heads, 64).permute(2, 0, 3, 1, 4).reshape(3, n_utt * heads, period, 64)
    assert range_err(q, projected[0] * (0.125 * 1.4426950408889634)) < 1e-2
    assert range_err(k, projected[1]) < 1e-2 and range_err(v, projected[2]) < 1e-2


@pytest.mark.parametrize("length_in,kernel,stride", [(1001, 3, 2), (3999, 3, 2), (999, 2, 2)])
def test_gemm_strided_conv(length_in, kernel, stride):
    """Implicit-GEMM Conv1d over a channels-last activation through an overlapping-row TMA view (HF:281-299)."""
    ops = _ops()
    torch.manual_seed(length_in)
    n, c = 2, 512
    x = (torch.randn(n, length_in, c, device=DEV) * 0.5).bfloat16()
    weight = (torch.randn(c, c, kernel, device=DEV) * 0.03).bfloat16()
    bias = torch.randn(c, device=DEV)
    length_out = (length_in - kernel) // stride + 1
    packed = ops.pack_conv_weight(weight.float())
    out = torch.zeros(n, length_out, c, device=DEV, dtype=torch.bfloat16)
    args = ops.make_gemm_args(
        x, packed, a_rows=length_out, a_inner=kernel * c, a_row_stride=stride * c, batch=n, a_batch_stride=length_in * c, bias=bias,
        out_bf16=out, ld_bf16=c, out_batch_rows=length_out,
    )  # fmt: skip
    ops.run_gemm(args)
    reference = F.conv1d(x.float().cpu().transpose(1, 2), weight.float().cpu(), bias.cpu(), stride=stride).transpose(1, 2)
    assert range_err(out, reference) < 1e-2


@pytest.mark.parametrize("frames", [499, 130, 17])
def test_gemm_positional_conv(frames):
    """Grouped 128-tap positional conv + GELU + residual (HF:326-368, 764-765) incl. weight_norm packing."""
    ops = _ops()
    torch.manual_seed(frames)
    n, c, groups, taps = 2, 1024, 16, 128
    x = (torch.randn(n, frames, c, device=DEV) * 0.5).bfloat16()
    v = torch.randn(c, c // groups, taps, device=DEV) * 0.02
    g = torch.rand(1, 1, taps, device=DEV) + 0.5
    bias = torch.randn(c, device=DEV)
    packed = ops.pack_posconv_weight(g, v)
    weight = (g * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()).cpu()  # weight_norm(dim=2)
    assert range_err(packed.float().view(c, taps, c // groups).permute(0, 2, 1), weight) < 1e-2
    hidden = x.float().contiguous()
    args = ops.make_gemm_args(
        x, packed, a_rows=frames, a_inner=c, a_row_stride=c, batch=n, a_batch_stride=frames * c, mode=1, tap_pad=taps // 2, n=c,
        k=taps * (c // groups), bias=bias, gelu=True, resid=hidden, ld_resid=c, out_f32=hidden, ld_f32=c, out_batch_rows=frames,
    )  # fmt: skip
    ops.run_gemm(args)
    conv = F.conv1d(x.float().cpu().transpose(1, 2), packed.float().cpu().view(c, taps, c // groups).permute(0, 2, 1).contiguous(),
                    bias.cpu(), padding=taps // 2, groups=groups)[:, :, :-1].transpose(1, 2)  # fmt: skip
    reference = x.float().cpu() + F.gelu(conv)
    assert range_err(hidden, reference) < 1e-4


def test_gemm_qkv_scatter_epilogue():
    """Fused Q/K/V projection: Q pre-scaled by head_dim^-0.5 * log2(e), K and V, all as bf16 [utterance*head, frame, 64]."""
    ops = _ops()
    torch.manual_seed(21)
    n_utt, seq, heads = 3, 249, 16
    hidden, rows = heads * 64, 3 * 249
    x = (torch.randn(rows, hidden, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(3 * hidden, hidden, device=DEV) * 0.03).bfloat16()
    bias = torch.randn(3 * hidden, device=DEV)
    q, k, v = (torch.full((n_utt * heads * seq * 64,), float("nan"), device=DEV, dtype=torch.bfloat16) for _ in range(3))
    ops.run_gemm(ops.make_qkv_args(x, w, bias, q, k, v, rows=rows, seq=seq, heads=heads))
    reference = (x.float() @ w.float().T + bias).view(n_utt, seq, 3, heads, 64).permute(2, 0, 3, 1, 4)  # [3, N, heads, T, 64]
    scale = 0.125 * math.log2(math.e)
    for ours, ref in ((q, reference[0] * scale), (k, reference[1]), (v, reference[2])):
        assert range_err(ours.view(n_utt, heads, seq, 64), ref) < 1e-2


# ------------------------------------------------------------------------------------------ attention
@pytest.fixture(params=[1, 2], ids=["blocks_of_64", "tile_pairs"])
def attention_kernel(request):
    """Both forward kernels on every shape (by default the library picks by problem size: ops.set_attention_kernel)."""
    ops = _ops()
    before = ops.set_attention_kernel(request.param)
    yield request.param
    ops.set_attention_kernel(before)


@pytest.mark.parametrize("n,heads,frames,lengths", [(1, 1, 128, [128]), (2, 16, 499, [499, 300]), (3, 4, 749, [749, 1, 130]), (1, 2, 1499, [1000])])
def test_attention(n, heads, frames, lengths, attention_kernel):
    ops = _ops()
    torch.manual_seed(frames)
    q = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
    k = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
    v = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
    q_scaled = (q.float() * 0.125 * math.log2(math.e)).bfloat16()  # the kernel works in the log2 domain
    ctx = torch.zeros(n * frames, heads * 64, device=DEV, dtype=torch.bfloat16)
    frame_lengths = torch.tensor(lengths, device=DEV, dtype=torch.int32)
    ops.attention(q_scaled, k, v, ctx, frame_lengths, n, heads, frames)
    mask = torch.arange(frames)[None, :] < torch.tensor(lengths)[:, None]
    scores = (q_scaled.float().cpu() / math.log2(math.e) @ k.float().cpu().transpose(2, 3)).masked_fill(~mask[:, None, None, :], float("-inf"))
    reference = (torch.softmax(scores, -1) @ v.float().cpu()).permute(0, 2, 1, 3).reshape(n, frames, heads * 64)
    ours = ctx.float().cpu().view(n, frames, heads * 64)
    for index, length in enumerate(lengths):
        assert range_err(ours[index, :length], reference[index, :length]) < 2e-2


def test_multi_head_block_kernels_match_torch():
    """One launch for all heads: column blocks of a level matrix copied out / accumulated back, and the per-head log_softmax in front
    of the CTC loss (time-first views keep their strides)."""
    ops = _ops()
    torch.manual_seed(3)
    rows, ld = 1237, 96
    level = torch.randn(rows, ld, device=DEV)
    layout = [(0, 5), (5, 31), (36, 2), (40, 47)]  # (column, width)
    outs = [torch.full((rows, width), float("nan"), device=DEV) for _, width in layout]
    ops.copy_head_blocks([(level, ld, column, width) for column, width in layout], [(out, width, 0, width) for out, (_, width) in zip(outs, layout)], rows)
    for out, (column, width) in zip(outs, layout):
        assert torch.equal(out, level[:, column : column + width])
    target = torch.randn(rows, ld, device=DEV)
    expected = target.clone()
    for out, (column, width) in zip(outs, layout):
        expected[:, column : column + width] += out
    ops.copy_head_blocks([(out, width, 0, width) for out, (_, width) in zip(outs, layout)], [(target, ld, column, width) for column, width in layout], rows, accumulate=True)
    assert torch.equal(target, expected)
    # blocks with their own row counts (weight blocks of different classifiers into one level matrix)
    big = torch.zeros(40, 64, device=DEV)
    parts = [torch.randn(3, 64, device=DEV), torch.randn(17, 20, device=DEV), torch.randn(1, 9, device=DEV)]
    ops.copy_head_blocks(
        [(parts[0], 64, 0, 64, 3), (parts[1], 20, 0, 20, 17), (parts[2], 9, 0, 9, 1)],
        [(big, 64, 0, 64, 3), (big, 64, 5 * 64 + 8, 20, 17), (big, 64, 30 * 64 + 50, 9, 1)], 0, accumulate=True,
    )  # fmt: skip
    reference = torch.zeros(40, 64, device=DEV)
    reference[:3] += parts[0]
    reference[5:22, 8:28] += parts[1]
    reference[30:31, 50:59] += parts[2]
    assert torch.equal(big, reference)
    n_utt, seq = 7, 53
    logits = [torch.randn(n_utt, seq, width, device=DEV).transpose(0, 1) * 3 for width in (4, 33, 501, 2)]  # time-first views
    logits.append(torch.randn(seq, n_utt, 9, device=DEV).transpose(0, 1).contiguous().transpose(0, 1))
    results = ops.log_softmax_many(logits)
    for ours, x in zip(results, logits):
        assert ours.shape == x.shape and ours.stride() == x.stride()
        assert float((ours - torch.log_softmax(x, -1)).abs().max()) < 2e-6
    mixed = ops.log_softmax_many([logits[0], torch.randn(3, 4, 5, device=DEV)])  # different row counts: one by one
    assert float((mixed[1] - torch.log_softmax(torch.randn(3, 4, 5, device=DEV) * 0 + mixed[1].exp().log(), -1)).abs().max()) < 1e-5


def test_attention_ragged_batches_on_the_persistent_schedule(attention_kernel):
    """Batches as a MaxFrameBatchSampler stream produces them (many utterances of very different lengths, frames padded to a
    multiple of 64): every persistent CTA works through a long list of items that mixes full query-tile pairs, pairs whose second
    tile is padding only (its MMA issuer still walks the ring barriers: skipping ahead aliases mbarrier parities and once hung the
    kernel) and pairs that are skipped altogether."""
    ops = _ops()
    generator = torch.Generator().manual_seed(5)
    heads = 4
    for _ in range(3):
        n = int(torch.randint(20, 60, (1,), generator=generator))
        seconds = torch.rand(n, generator=generator) * 12 + 3
        lengths = (((seconds * 16000 - 400) / 320).floor().int() + 1).tolist()
        lengths[0] = 1
        lengths[1] = 129
        frames = (max(lengths) + 63) // 64 * 64
        q = torch.randn(n, heads, frames, 64, device=DEV, generator=None).bfloat16()
        k = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
        v = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
        q_scaled = (q.float() * 0.125 * math.log2(math.e)).bfloat16()
        ctx = torch.full((n * frames, heads * 64), float("nan"), device=DEV, dtype=torch.bfloat16)
        frame_lengths = torch.tensor(lengths, device=DEV, dtype=torch.int32)
        ops.attention(q_scaled, k, v, ctx, frame_lengths, n, heads, frames)
        mask = torch.arange(frames, device=DEV)[None, :] < frame_lengths[:, None]
        scores = (q_scaled.float() / math.log2(math.e) @ k.float().transpose(2, 3)).masked_fill(~mask[:, None, None, :], float("-inf"))
        reference = (torch.softmax(scores, -1) @ v.float()).permute(0, 2, 1, 3).reshape(n, frames, heads * 64)
        ours = ctx.float().view(n, frames, heads * 64)
        assert torch.isfinite(ours).all()  # rows of padded queries are written too (zeros or values computed from real keys)
        for index, length in enumerate(lengths):
            assert range_err(ours[index, :length], reference[index, :length]) < 2e-2


@pytest.mark.parametrize("sharpness", [8.0, 40.0])
def test_attention_large_scores_exercise_lazy_rescaling(sharpness, attention_kernel):
    """Peaked score distributions (row maxima that keep growing by more than 2^8 from block to block) exercise the
    in-TMEM rescaling of the running output and its barrier protocol; many CTAs run concurrently (warps of one CTA
    drift apart by a block), repeated launches must agree bit for bit."""
    ops = _ops()
    torch.manual_seed(7)
    n, heads, frames = 8, 16, 499
    lengths = [499, 450, 400, 333, 257, 129, 64, 17]
    q = torch.randn(n, heads, frames, 64, device=DEV)
    k = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
    v = torch.randn(n, heads, frames, 64, device=DEV).bfloat16()
    # later keys get larger scores: the running maximum grows along the key axis
    k = (k.float() * (1.0 + torch.arange(frames, device=DEV)[None, None, :, None] / frames)).bfloat16()
    q_scaled = (q * sharpness * 0.125 * math.log2(math.e)).bfloat16()
    frame_lengths = torch.tensor(lengths, device=DEV, dtype=torch.int32)
    results = []
    for _ in range(3):
        ctx = torch.zeros(n * frames, heads * 64, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(n * heads * frames, device=DEV, dtype=torch.float32)
        ops.attention(q_scaled, k, v, ctx, frame_lengths, n, heads, frames, lse)
        results.append((ctx, lse))
    torch.cuda.synchronize()
    mask = torch.arange(frames)[None, :] < torch.tensor(lengths)[:, None]
    scores = (q_scaled.float().cpu() / math.log2(math.e) @ k.float().cpu().transpose(2, 3)).masked_fill(~mask[:, None, None, :], float("-inf"))
    reference = (torch.softmax(scores, -1) @ v.float().cpu()).permute(0, 2, 1, 3).reshape(n, frames, heads * 64)
    lse_ref = torch.logsumexp(scores, -1) * math.log2(math.e)
    ours = results[0][0].float().cpu().view(n, frames, heads * 64)
    lse_ours = results[0][1].view(n, heads, frames).cpu()
    for index, length in enumerate(lengths):
        assert range_err(ours[index, :length], reference[index, :length]) < 2e-2
        assert float((lse_ours[index, :, :length] - lse_ref[index, :, :length]).abs().max()) < 2e-2 * max(1.0, float(lse_ref[index, :, :length].abs().max()))
    for ctx, lse in results[1:]:
        assert torch.equal(ctx, results[0][0]) and torch.equal(lse, results[0][1])


# ------------------------------------------------------------------------------------------ front end
def test_wave_norm_and_frame_lengths():
    ops = _ops()
    lengths = torch.tensor([16000, 12345, 400, 10])
    audio = restatement.synthetic_audio(4, 16000, seed=3) + 0.05
    audio = audio * restatement.mask_sequence(lengths)
    reference = restatement.zero_mean_unit_var_norm(audio, lengths, restatement.mask_sequence(lengths))
    ours = ops.zero_mean_unit_var_norm(audio.cuda(), lengths.cuda())
    assert range_err(ours, reference) < 1e-5
    assert torch.equal(ours.cpu()[1, 12345:], torch.zeros(16000 - 12345))
    kernels = torch.tensor(restatement.XLSR_300M["conv_kernel"], dtype=torch.int32, device=DEV)
    strides = torch.tensor(restatement.XLSR_300M["conv_stride"], dtype=torch.int32, device=DEV)
    many = torch.tensor([16000, 12345, 400, 10, 480000, 399, 401, 80000, 160000, 240000, 5, 0])
    frames32 = torch.empty(len(many), dtype=torch.int32, device=DEV)
    frames64 = torch.empty(len(many), dtype=torch.int64, device=DEV)
    ops.frame_lengths(many.cuda(), kernels, strides, frames32, frames64)
    expected = restatement.conv_lengths(many, restatement.XLSR_300M["conv_kernel"], restatement.XLSR_300M["conv_stride"])
    assert torch.equal(frames64.cpu(), expected)
    assert torch.equal(frames32.cpu().long(), expected)


def test_conv0_layernorm_gelu():
    ops = _ops()
    torch.manual_seed(5)
    lengths = torch.tensor([8000, 5003])
    audio = restatement.synthetic_audio(2, 8000, seed=4) * restatement.mask_sequence(lengths)
    weight = torch.randn(512, 1, 10) * 0.3
    bias, gamma, beta = torch.randn(512) * 0.1, torch.rand(512) + 0.5, torch.randn(512) * 0.1
    normalised = restatement.zero_mean_unit_var_norm(audio, lengths, restatement.mask_sequence(lengths))
    conv = F.conv1d(normalised[:, None], weight, bias, stride=5).transpose(1, 2)
    reference = F.gelu(F.layer_norm(conv, (512,), gamma, beta, 1e-5))
    frames0 = (8000 - 10) // 5 + 1
    stats = torch.empty(2, 3, dtype=torch.float64, device=DEV)
    mean_rstd = torch.empty(2, 2, device=DEV)
    ops.wave_stats(audio.cuda(), lengths.cuda(), stats, mean_rstd)
    out = torch.zeros(2, frames0, 512, device=DEV, dtype=torch.bfloat16)
    ops.conv0_ln_gelu(audio.cuda(), lengths.cuda(), mean_rstd, weight.view(512, 10).cuda(), bias.cuda(), gamma.cuda(), beta.cuda(), 1e-5, out)
    valid = [(int(l) - 10) // 5 + 1 for l in lengths]
    for index, frames in enumerate(valid):
        assert range_err(out[index, :frames], reference[index, :frames]) < 1e-2
    # frames that would read padding are skipped (buffer stays zero; a warp finishes its group of 4 frames) ...
    assert float(out[1, (valid[1] + 3) // 4 * 4 :].float().abs().max()) == 0.0
    # ... unless asked not to
    ops.conv0_ln_gelu(audio.cuda(), lengths.cuda(), mean_rstd, weight.view(512, 10).cuda(), bias.cuda(), gamma.cuda(), beta.cuda(), 1e-5, out, False)
    assert range_err(out, reference) < 1e-2


def test_conv0_groupnorm_gelu():
    """feat_extract_norm="group" (wav2vec2-base style, HF:302-323): per-channel statistics over the padded time axis."""
    ops = _ops()
    torch.manual_seed(6)
    lengths = torch.tensor([6000, 4000])
    audio = restatement.synthetic_audio(2, 6000, seed=7) * restatement.mask_sequence(lengths)
    weight = torch.randn(512, 1, 10) * 0.3
    gamma, beta = torch.rand(512) + 0.5, torch.randn(512) * 0.1
    normalised = restatement.zero_mean_unit_var_norm(audio, lengths, restatement.mask_sequence(lengths))
    conv = F.conv1d(normalised[:, None], weight, None, stride=5)
    reference = F.gelu(F.group_norm(conv, 512, gamma, beta, 1e-5)).transpose(1, 2)
    frames0 = conv.shape[-1]
    stats = torch.empty(2, 3, dtype=torch.float64, device=DEV)
    mean_rstd = torch.empty(2, 2, device=DEV)
    ops.wave_stats(audio.cuda(), lengths.cuda(), stats, mean_rstd)
    raw = torch.empty(2, frames0, 512, device=DEV)
    gn_stats = torch.empty(2, 512, 2, dtype=torch.float64, device=DEV)
    out = torch.zeros(2, frames0, 512, device=DEV, dtype=torch.bfloat16)
    ops.conv0_gn_gelu(audio.cuda(), lengths.cuda(), mean_rstd, weight.view(512, 10).cuda(), None, gamma.cuda(), beta.cuda(), 1e-5, raw, gn_stats, out)
    assert range_err(out, reference) < 1e-2


@pytest.mark.parametrize("cols,dtype", [(512, torch.bfloat16), (1024, torch.float32), (1024, torch.bfloat16), (512, torch.float32)])
def test_layernorm_rows(cols, dtype):
    ops = _ops()
    torch.manual_seed(cols)
    rows = 1003
    x = (torch.randn(rows, cols) * 2 + 0.3).to(dtype)
    gamma, beta = torch.rand(cols) + 0.5, torch.randn(cols) * 0.1
    reference = F.layer_norm(x.float(), (cols,), gamma, beta, 1e-5)
    out32 = torch.empty(rows, cols, device=DEV)
    out16 = torch.empty(rows, cols + 64, device=DEV, dtype=torch.bfloat16)
    ops.layernorm_rows(x.cuda(), rows, cols, cols, gamma.cuda(), beta.cuda(), 1e-5, out_bf16=out16, ld_bf16=cols + 64, out_f32=out32, ld_f32=cols)
    assert range_err(out32, reference) < 1e-5
    assert range_err(out16[:, :cols], reference) < 1e-2
    ops.layernorm_rows(x.cuda(), rows, cols, cols, gamma.cuda(), beta.cuda(), 1e-5, gelu=True, out_f32=out32, ld_f32=cols)
    assert range_err(out32, F.gelu(reference)) < 1e-5


# ------------------------------------------------------------------------------------------ heads
def test_compose_embeddings_matches_embedding_bag():
    ops = _ops()
    torch.manual_seed(8)
    spec = restatement.multitask_spec(n_train_phonemes=40)
    oracle_table = torch.from_numpy(spec.feature_table).long()
    num_categories = torch.cat((torch.LongTensor([0]), oracle_table.max(0).values)) + 1
    offsets = num_categories.cumsum(0)[:-1]
    bag = torch.nn.EmbeddingBag(int(num_categories.sum()), 640, mode="sum")
    tfi = torch.randint(0, 3, (25, 36))
    reference = torch.cat((bag(torch.zeros(1, 1, dtype=torch.long)), bag(tfi + offsets))).detach()
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    out32 = torch.empty(32, 640, device=DEV)
    out16 = torch.empty(32, 640, device=DEV, dtype=torch.bfloat16)
    ops.compose_embeddings(bag.weight.detach().cuda(), tfi.cuda(), offsets.cuda(), 32, err, out_bf16=out16, out_f32=out32)
    assert int(err.item()) == 0
    assert range_err(out32[:26], reference) < 1e-6
    assert float(out32[26:].abs().max()) == 0.0
    assert range_err(out16[:26], reference) < 1e-2
    ops.compose_embeddings(bag.weight.detach().cuda(), (tfi + 5).cuda(), offsets.cuda(), 32, err, out_f32=out32)
    assert int(err.item()) == 1  # out-of-range category is flagged, not read


def test_log_softmax_heads_and_wide():
    ops = _ops()
    torch.manual_seed(9)
    rows = 1000
    widths = [4, 4, 3, 7, 26, 2, 64]
    columns, position = [], 8
    for width in widths:
        columns.append(position)
        position += width
    ld = (position + 3) // 4 * 4
    logits = torch.randn(rows, ld) * 3
    logits[5, columns[0] : columns[0] + 4] = 1.25  # exact tie -> lowest index wins
    out_offsets, total = [], 0
    for width in widths:
        out_offsets.append(total)
        total += rows * width
    out = torch.empty(total, device=DEV)
    argmax = torch.empty(len(widths), rows, dtype=torch.int32, device=DEV)
    maxlp = torch.empty(len(widths), rows, device=DEV)
    ops.log_softmax_heads(
        logits.cuda(), ld, rows, 8, ld - 8,
        torch.tensor(columns, dtype=torch.int32, device=DEV), torch.tensor(widths, dtype=torch.int32, device=DEV),
        torch.tensor(out_offsets, dtype=torch.int64, device=DEV), len(widths), out, argmax, maxlp,
    )  # fmt: skip
    for head, (column, width) in enumerate(zip(columns, widths)):
        reference = F.log_softmax(logits[:, column : column + width], -1)
        ours = out[out_offsets[head] : out_offsets[head] + rows * width].view(rows, width).cpu()
        assert float((ours - reference).abs().max()) < 1e-5
        values, indices = reference.max(-1)
        assert torch.equal(argmax[head].cpu().long(), indices)
        assert float((maxlp[head].cpu() - values).abs().max()) < 1e-5
    assert int(argmax[0, 5]) == 0
    # wide head
    wide = torch.randn(300, 3184) * 4
    out_wide = torch.empty(300, 3184, device=DEV)
    arg_wide = torch.empty(300, dtype=torch.int32, device=DEV)
    max_wide = torch.empty(300, device=DEV)
    ops.log_softmax_wide(wide.cuda(), 3184, 300, 3184, out_wide, 3184, arg_wide, max_wide)
    reference = F.log_softmax(wide, -1)
    assert float((out_wide.cpu() - reference).abs().max()) < 1e-4
    assert torch.equal(arg_wide.cpu().long(), reference.argmax(-1))
    assert float(torch.logsumexp(out_wide, -1).abs().max()) < 1e-4  # rows are normalised
    assert float((ops.log_softmax(wide.cuda().view(3, 100, 3184)).cpu() - reference.view(3, 100, 3184)).abs().max()) < 1e-4


def test_dependency_softmax():
    ops = _ops()
    torch.manual_seed(10)
    rows, ld = 500, 160
    logits = torch.randn(rows, ld) * 2
    columns, widths, targets = [0, 4, 100], [4, 4, 27], [1024, 1028, 1040]
    for skip in (0, 1):
        dst = torch.zeros(rows, 1088, dtype=torch.bfloat16, device=DEV)
        ops.dependency_softmax(
            logits.cuda(), ld, rows, torch.tensor(columns, dtype=torch.int32, device=DEV), torch.tensor(widths, dtype=torch.int32, device=DEV),
            torch.tensor(targets, dtype=torch.int32, device=DEV), 3, skip, dst, 1088,
        )  # fmt: skip
        for column, width, target in zip(columns, widths, targets):
            reference = torch.softmax(logits[:, column + skip : column + width], -1)
            assert range_err(dst[:, target : target + width - skip], reference) < 1e-2


# ------------------------------------------------------------------------------------------ greedy decoding
def test_greedy_decode_matches_oracle_exactly():
    from allophant_b200.predictions import GreedyCTCDecoder

    torch.manual_seed(11)
    n, frames, classes = 6, 211, 5
    emissions = F.log_softmax(torch.randn(n, frames, classes) * 2, -1)
    emissions[0, :, 1:] -= 50  # all blank
    emissions[1, :, 0] -= 50  # never blank
    emissions[2, 10:40] = emissions[2, 10]  # one long run
    lengths = torch.tensor([211, 200, 150, 1, 0, 33])
    reference = restatement.greedy_ctc_decode(emissions, lengths)
    ours = GreedyCTCDecoder()(emissions.cuda(), lengths.cuda())
    for hypothesis, expected in zip(ours, reference):
        assert torch.equal(hypothesis[0].tokens, expected[0].tokens)
        assert torch.equal(hypothesis[0].timesteps, expected[0].timesteps)
        assert abs(float(hypothesis[0].score) - float(expected[0].score)) < 1e-3 * max(1.0, abs(float(expected[0].score)))
        assert hypothesis[0].words == []
    # wide emissions use the warp-per-frame argmax
    wide = F.log_softmax(torch.randn(2, 90, 300), -1)
    lengths = torch.tensor([90, 45])
    for hypothesis, expected in zip(GreedyCTCDecoder()(wide.cuda(), lengths.cuda()), restatement.greedy_ctc_decode(wide, lengths)):
        assert torch.equal(hypothesis[0].tokens, expected[0].tokens)
        assert torch.equal(hypothesis[0].timesteps, expected[0].timesteps)


# ------------------------------------------------------------------------------------------ CTC
def _ctc_case(seed, n, frames, class_counts, label_fraction, time_first=True):
    generator = torch.Generator().manual_seed(seed)
    input_lengths = torch.randint(frames // 2, frames + 1, (n,), generator=generator)
    input_lengths[0] = frames
    logits, labels, label_lengths = [], [], []
    for head, classes in enumerate(class_counts):
        shape = (frames, n, classes) if time_first else (n, frames, classes)
        logits.append(torch.randn(*shape, generator=generator) * 2)
        head_labels, head_lengths = restatement.synthetic_labels(input_lengths, classes, seed=seed * 31 + head, fraction=label_fraction)
        labels.append(head_labels)
        label_lengths.append(head_lengths)
    return logits, labels, input_lengths, label_lengths


@pytest.mark.parametrize("class_counts,fraction", [([4, 4, 3, 26], 0.25), ([4], 0.5), ([61, 500], 0.3), ([4, 40], 0.02)])
def test_ctc_loss_and_gradient_match_torch(class_counts, fraction):
    from allophant_b200.loss_functions import CTCWrapper, multi_head_ctc_loss

    logits, labels, input_lengths, label_lengths = _ctc_case(len(class_counts), 5, 120, class_counts, fraction)
    # make one utterance infeasible (labels longer than frames): zero_infinity must zero loss AND gradient
    label_lengths[0][1] = min(int(labels[0].shape[1]), int(input_lengths[1]) + 5)
    if label_lengths[0][1] <= input_lengths[1]:
        input_lengths[1] = max(1, int(label_lengths[0][1]) - 1)
    # repeated labels need an extra blank between them
    labels[0][2, : max(1, int(label_lengths[0][2]))] = 1
    reference_losses, reference_grads = [], []
    for head_logits, head_labels, head_lengths in zip(logits, labels, label_lengths):
        leaf = head_logits.clone().requires_grad_(True)
        loss = restatement.ctc_wrapper(leaf, head_labels, input_lengths, head_lengths)
        loss.backward()
        reference_losses.append(float(loss))
        reference_grads.append(leaf.grad)
    leaves = [t.cuda().requires_grad_(True) for t in logits]
    losses = multi_head_ctc_loss(leaves, [l.cuda() for l in labels], input_lengths.cuda(), [l.cuda() for l in label_lengths])
    weights = torch.arange(1, len(class_counts) + 1, device=DEV, dtype=torch.float32)
    (losses * weights).sum().backward()
    for head in range(len(class_counts)):
        assert abs(float(losses[head]) - reference_losses[head]) <= 1e-3 * max(1.0, abs(reference_losses[head]))
        ours = leaves[head].grad.cpu() / float(weights[head])
        scale = float(reference_grads[head].abs().max().clamp_min(1e-6))
        assert float((ours - reference_grads[head]).abs().max()) <= 1e-3 * scale + 1e-5
    # the drop-in single-head wrapper
    single = CTCWrapper()(logits[0].cuda(), labels[0].cuda(), input_lengths.cuda(), label_lengths[0].cuda())
    assert abs(float(single) - reference_losses[0]) <= 1e-3 * max(1.0, abs(reference_losses[0]))


def test_ctc_long_labels_and_empty_targets():
    from allophant_b200.loss_functions import multi_head_ctc_loss

    logits, labels, input_lengths, label_lengths = _ctc_case(21, 3, 700, [4, 30], 0.45)
    label_lengths[1][2] = 0  # empty target
    leaves = [t.cuda().requires_grad_(True) for t in logits]
    losses = multi_head_ctc_loss(leaves, [l.cuda() for l in labels], input_lengths.cuda(), [l.cuda() for l in label_lengths])
    losses.sum().backward()
    for head in range(2):
        leaf = logits[head].clone().requires_grad_(True)
        reference = restatement.ctc_wrapper(leaf, labels[head], input_lengths, label_lengths[head])
        reference.backward()
        # 700-frame log-space recursions in fp32 accumulate rounding (alpha ~ -1e3): judge both fp32
        # implementations against the same recursion in fp64
        leaf64 = logits[head].double().requires_grad_(True)
        exact = restatement.ctc_wrapper(leaf64, labels[head], input_lengths, label_lengths[head])
        exact.backward()
        assert abs(float(losses[head]) - float(exact)) <= 1e-3 * abs(float(exact))
        scale = float(leaf64.grad.abs().max())
        ours_error = float((leaves[head].grad.cpu().double() - leaf64.grad).abs().max())
        torch_error = float((leaf.grad.double() - leaf64.grad).abs().max())
        print(f"ctc long head {head}: ours vs fp64 {ours_error:.2e}, torch fp32 vs fp64 {torch_error:.2e}")
        assert ours_error <= max(3 * torch_error, 1e-3 * scale)
        # size-independent property: on valid frames the gradient rows of a finite loss sum to zero
        row_sums = leaves[head].grad.sum(-1).cpu()
        assert float(row_sums.abs().max()) < 5e-3  # |nll| ~ 1e3 in fp32: 1e-6 relative on nll = 1e-3 on sum(gamma)


def test_ctc_nan_losses_stay_visible_and_bad_labels_poison_the_loss():
    """zero_infinity zeroes INFINITE losses only (loss_functions.py:24): a NaN loss (diverged logits) stays NaN in the per-head sum and
    in the gradient of its utterance, as in torch; a label outside the head's classes (nn.CTCLoss raises) gives a NaN loss instead of
    indexing the staged log-probabilities out of bounds."""
    from allophant_b200.loss_functions import multi_head_ctc_loss

    logits, labels, input_lengths, label_lengths = _ctc_case(5, 3, 120, [6, 40], 0.3)
    logits[0][10, 1, 2] = float("nan")  # utterance 1 of the narrow head diverged
    labels[1][2, 0] = 40  # utterance 2 of the wide head: label == number of classes
    leaves = [t.cuda().requires_grad_(True) for t in logits]
    losses = multi_head_ctc_loss(leaves, [l.cuda() for l in labels], input_lengths.cuda(), [l.cuda() for l in label_lengths])
    assert torch.isnan(losses).all()
    losses.sum().backward()
    for head, poisoned in ((0, 1), (1, 2)):
        grad = leaves[head].grad.cpu()
        frames = int(input_lengths[poisoned])
        assert torch.isnan(grad[:frames, poisoned]).all()
        assert float(grad[frames:, poisoned].abs().max()) == 0.0 if frames < grad.shape[0] else True
        clean = [n for n in range(3) if n != poisoned]
        assert torch.isfinite(grad[:, clean]).all()  # the other utterances of the head are untouched


def test_ctc_label_sequences_beyond_511():
    """More than 511 labels (2S+1 states no longer fit a warp's registers): the block-per-pair path, against the fp64 recursion.
    nn.CTCLoss has no cap (loss_functions.py:24); contour attributes and 30 s utterances can exceed 511 labels."""
    from allophant_b200.loss_functions import multi_head_ctc_loss

    torch.manual_seed(5)
    n_utt, frames = 3, 1499
    class_counts = [4, 40]
    input_lengths = torch.tensor([1499, 1300, 900])
    label_lengths = [torch.tensor([700, 520, 0]), torch.tensor([600, 3, 450])]
    logits, labels = [], []
    for classes, lengths in zip(class_counts, label_lengths):
        logits.append(torch.randn(frames, n_utt, classes) * 2)  # time-first, like the model's outputs
        head_labels = torch.randint(1, classes, (n_utt, int(max(lengths))))
        head_labels[0, 10:14] = 1  # repeats need blanks between them
        labels.append(head_labels)
    leaves = [t.cuda().requires_grad_(True) for t in logits]
    losses = multi_head_ctc_loss(leaves, [l.cuda() for l in labels], input_lengths.cuda(), [l.cuda() for l in label_lengths])
    weights = torch.tensor([1.0, 0.5], device=DEV)
    (losses * weights).sum().backward()
    for head in range(2):
        leaf64 = logits[head].double().requires_grad_(True)
        exact = restatement.ctc_wrapper(leaf64, labels[head], input_lengths, label_lengths[head])
        exact.backward()
        assert abs(float(losses[head]) - float(exact)) <= 1e-3 * abs(float(exact)), (float(losses[head]), float(exact))
        ours = leaves[head].grad.cpu().double() / float(weights[head])
        scale = float(leaf64.grad.abs().max())
        # 1 499-frame log-space recursions in fp32 (|alpha| ~ 4e3, 2.4e-4 absolute per operation): judged like the 700-frame case,
        # against what torch's own fp32 CTC loses to the fp64 recursion
        leaf32 = logits[head].clone().requires_grad_(True)
        restatement.ctc_wrapper(leaf32, labels[head], input_lengths, label_lengths[head]).backward()
        torch_error = float((leaf32.grad.double() - leaf64.grad).abs().max())
        ours_error = float((ours - leaf64.grad).abs().max())
        print(f"ctc beyond 511 labels, head {head}: ours vs fp64 {ours_error:.2e}, torch fp32 vs fp64 {torch_error:.2e}")
        assert ours_error <= max(3 * torch_error, 2e-3 * scale)
        assert float(ours[900:, 2].abs().max()) == 0.0  # frames past the utterance


def test_custom_ops_run_the_library_kernels():
    """torch.ops.allophant_b200.* against torch's own ops (and autograd through log_softmax)."""
    import allophant_b200.custom_ops  # noqa: F401  (registers the ops)

    torch.manual_seed(5)
    x = torch.randn(50, 3, 40, device=DEV, requires_grad=True)
    out = torch.ops.allophant_b200.log_softmax(x)
    assert range_err(out, F.log_softmax(x.detach(), -1)) < 1e-5
    grad = torch.randn_like(out)
    out.backward(grad)
    reference = x.detach().clone().requires_grad_(True)
    F.log_softmax(reference, -1).backward(grad)
    assert range_err(x.grad, reference.grad) < 1e-5
    weight, bias = torch.randn(64, 40, device=DEV) * 0.1, torch.randn(64, device=DEV)
    linear = torch.ops.allophant_b200.linear_bf16(x.detach(), weight, bias, False)
    assert range_err(linear, F.linear(x.detach().bfloat16().float(), weight.bfloat16().float(), bias)) < 1e-4
    wide = torch.randn(33, 1024, device=DEV)
    gamma, beta = torch.randn(1024, device=DEV), torch.randn(1024, device=DEV)
    assert range_err(torch.ops.allophant_b200.layer_norm(wide, gamma, beta, 1e-5), F.layer_norm(wide, (1024,), gamma, beta, 1e-5)) < 1e-5
    labels = torch.randint(1, 40, (3, 9), device=DEV)
    lengths, label_lengths = torch.tensor([50, 40, 30], device=DEV), torch.tensor([9, 5, 1], device=DEV)
    nll = torch.ops.allophant_b200.ctc_nll(out.detach(), labels, lengths, label_lengths)
    reference = F.ctc_loss(out.detach().cpu(), labels.cpu(), lengths.cpu(), label_lengths.cpu(), reduction="none")
    assert range_err(nll, reference) < 1e-4
    # CTC loss on the logits, differentiable (CTCWrapper.forward on one head)
    from allophant_b200.custom_ops import ctc_loss

    logits = x.detach().clone().requires_grad_(True)
    loss = ctc_loss(logits, labels, lengths, label_lengths)
    loss.backward()
    leaf = x.detach().clone().requires_grad_(True)
    expected = F.ctc_loss(F.log_softmax(leaf, -1), labels, lengths, label_lengths, reduction="sum", zero_infinity=True)
    expected.backward()
    assert abs(float(loss) - float(expected)) < 1e-3 * abs(float(expected))
    assert range_err(logits.grad, leaf.grad) < 1e-4
    # greedy decode = collapse of the frame-wise argmax
    tokens, counts, _ = torch.ops.allophant_b200.ctc_greedy_decode(out.detach(), lengths)
    best = out.detach().argmax(-1).transpose(0, 1).cpu()
    for row in range(3):
        frames = best[row, : int(lengths[row])].tolist()
        collapsed = [c for i, c in enumerate(frames) if c != 0 and (i == 0 or c != frames[i - 1])]
        assert tokens[row, : int(counts[row])].tolist() == collapsed
    # attention against SDPA with the key-padding mask
    heads, seq = 2, 200
    q, k, v = (torch.randn(3 * heads, seq, 64, device=DEV).bfloat16() for _ in range(3))
    frame_lengths = torch.tensor([200, 130, 1], device=DEV, dtype=torch.int32)
    ctx = torch.ops.allophant_b200.attention(q, k, v, frame_lengths, heads)
    mask = (torch.arange(seq, device=DEV)[None, :] < frame_lengths[:, None])[:, None, None, :]
    expected_ctx = F.scaled_dot_product_attention(q.float().view(3, heads, seq, 64), k.float().view(3, heads, seq, 64), v.float().view(3, heads, seq, 64), attn_mask=mask)
    expected_ctx = expected_ctx.transpose(1, 2).reshape(3 * seq, heads * 64)
    valid = (torch.arange(seq, device=DEV)[None, :] < frame_lengths[:, None]).reshape(-1)
    assert range_err(ctx.float()[valid], expected_ctx[valid]) < 2e-2
