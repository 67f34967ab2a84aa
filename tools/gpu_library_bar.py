"""Same-box library bar ("beat the library kernel torch dispatches on the same B200, measured in the same run").

Times, in ONE process on one B200, every hot kernel of this repository next to the library kernel torch would dispatch for the
same arithmetic on the same shapes:

* the four encoder linears of BASELINE configs[1] (M = 32 x 499 = 15 968 rows): cuBLASLt bf16 through ``torch.matmul`` (+ the
  separate bias / GELU / residual passes the reference's eager graph runs) against ``aph_gemm_bf16`` with its fused epilogues;
* self-attention 32 x 16 heads x 499 frames x 64: ``torch.nn.functional.scaled_dot_product_attention`` (flash backend) against
  ``aph_attention_bf16`` (which also handles the key-padding from frame counts);
* ``log_softmax`` on the config-3 wide head (63 872 x 3 184) and ``F.ctc_loss`` on the config-2 sizes against ``aph_log_softmax_wide``
  and ``aph_ctc_forward/backward``;
* LayerNorm 15 968 x 1 024 fp32 -> bf16.

Inputs are larger than L2 or rotated over several buffers so that no timing runs out of a warm cache.  Prints a markdown table.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from allophant_b200 import ops  # noqa: E402

dev = "cuda"
ROTATE = 4  # independent operand sets per timing: 4 x (>= 65 MB) exceeds the 126 MB L2


def timeit(calls, iters=12, warm=3):
    """Average milliseconds per call over ``iters`` passes of the list ``calls`` (one entry per rotated buffer set)."""
    for _ in range(warm):
        for call in calls:
            call()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        for call in calls:
            call()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / (iters * len(calls))


rows = []


def row(name, ours_ms, lib_ms, work, unit, note=""):
    """``work`` = FLOPs (unit "TFLOP/s") or bytes (unit "GB/s") of one call."""
    per_ms = 1e9 if unit == "TFLOP/s" else 1e6
    rows.append(dict(name=name, ours_us=ours_ms * 1e3, library_us=lib_ms * 1e3, ours_rate=work / ours_ms / per_ms, library_rate=work / lib_ms / per_ms,
                     unit=unit, speedup=lib_ms / ours_ms, note=note))  # fmt: skip


M = 32 * 499
peaks_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# ---------------------------------------------------------------------------------------------- encoder linears
for name, n, k, epilogue in (("QKV 1024->3072 (+bias)", 3072, 1024, "bias"), ("out-proj 1024->1024 (+bias +fp32 residual)", 1024, 1024, "resid"),
                             ("FFN1 1024->4096 (+bias +GELU)", 4096, 1024, "gelu"), ("FFN2 4096->1024 (+bias +fp32 residual)", 1024, 4096, "resid")):  # fmt: skip
    xs = [torch.randn(M, k, device=dev).bfloat16() for _ in range(ROTATE)]
    ws = [(torch.randn(n, k, device=dev) * 0.03).bfloat16() for _ in range(ROTATE)]
    bias = torch.randn(n, device=dev)
    bias16 = bias.bfloat16()
    resid = torch.randn(M, n, device=dev) if epilogue == "resid" else None
    outs16 = [torch.empty(M, n, device=dev, dtype=torch.bfloat16) for _ in range(ROTATE)]
    outs32 = [torch.randn(M, n, device=dev) for _ in range(ROTATE)] if epilogue == "resid" else None  # the residual stream, updated in place
    flops = 2.0 * M * n * k

    # library, GEMM only (the bar for the tensor pipe) and the eager graph of the reference (GEMM + separate epilogue passes)
    plain = timeit([lambda x=x, w=w, o=o: torch.matmul(x, w.t(), out=o) for x, w, o in zip(xs, ws, outs16)])

    def eager(x, w):
        y = torch.nn.functional.linear(x, w, bias16)
        if epilogue == "gelu":
            y = torch.nn.functional.gelu(y)
        if epilogue == "resid":
            y = resid + y.float()
        return y

    fused_lib = timeit([lambda x=x, w=w: eager(x, w) for x, w in zip(xs, ws)])
    calls = []
    for i in range(ROTATE):
        if epilogue == "resid":
            # as the encoder's launch list does it: hidden += Linear(x), residual read and result written in place
            args = ops.make_gemm_args(xs[i], ws[i], a_rows=M, a_inner=k, a_row_stride=k, bias=bias, resid=outs32[i], ld_resid=n, out_f32=outs32[i], ld_f32=n)
        else:
            args = ops.make_gemm_args(xs[i], ws[i], a_rows=M, a_inner=k, a_row_stride=k, bias=bias, gelu=epilogue == "gelu", out_bf16=outs16[i], ld_bf16=n)
        calls.append(lambda args=args: ops.run_gemm(args))
    ours = timeit(calls)
    # numerics of the comparison itself
    reference = eager(xs[-1], ws[-1]).float()
    if epilogue == "resid":
        outs32[-1].copy_(resid)
        calls[-1]()
    mine = (outs32[-1] if epilogue == "resid" else outs16[-1]).float()
    deviation = float((mine - reference).abs().max() / reference.abs().max())
    # the bar is the same ARITHMETIC: what torch launches for this Linear (+ GELU / + fp32 residual) in the reference's eager graph;
    # the cuBLASLt GEMM alone (bf16 out, no epilogue) is quoted next to it as the tensor-pipe bar
    row(f"GEMM {name}", ours, fused_lib, flops, "TFLOP/s",
        f"library = F.linear (cuBLASLt, bias fused) + its eager GELU / residual passes; the cuBLASLt GEMM alone: {plain * 1e3:.1f} us = "
        f"{flops / plain / 1e9:.0f} TFLOP/s ({plain / ours:.2f}x of ours); max dev {deviation:.1e}")
    del xs, ws, outs16, outs32, resid

# ---------------------------------------------------------------------------------------------- attention
N, H, T, D = 32, 16, 499, 64
qs = [torch.randn(N * H, T, D, device=dev).bfloat16() for _ in range(ROTATE)]
ks = [torch.randn(N * H, T, D, device=dev).bfloat16() for _ in range(ROTATE)]
vs = [torch.randn(N * H, T, D, device=dev).bfloat16() for _ in range(ROTATE)]
ctx = torch.empty(N * T, H * D, device=dev, dtype=torch.bfloat16)
lengths = torch.full((N,), T, dtype=torch.int32, device=dev)
att_flops = 4.0 * N * H * T * T * D
ours = timeit([lambda q=q, k=k, v=v: ops.attention(q, k, v, ctx, lengths, N, H, T) for q, k, v in zip(qs, ks, vs)])
q4 = [q.view(N, H, T, D) for q in qs]
k4 = [k.view(N, H, T, D) for k in ks]
v4 = [v.view(N, H, T, D) for v in vs]
sdpa = timeit([lambda q=q, k=k, v=v: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0) for q, k, v in zip(q4, k4, v4)])
row("attention 32 x 16 x 499 x 64 (no padding)", ours, sdpa, att_flops, "TFLOP/s", "library = torch SDPA (flash), no mask; ours applies the key-padding from frame counts")
ragged = torch.randint(150, T + 1, (N,), dtype=torch.int32, device=dev)
ragged[0] = T
ours_r = timeit([lambda q=q, k=k, v=v: ops.attention(q, k, v, ctx, ragged, N, H, T) for q, k, v in zip(qs, ks, vs)])
mask = (torch.arange(T, device=dev)[None, :] < ragged[:, None])[:, None, None, :]
sdpa_r = timeit([lambda q=q, k=k, v=v: torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=1.0) for q, k, v in zip(q4, k4, v4)])
useful = 4.0 * float((ragged.double() ** 2).sum()) * H * D
row("attention, ragged lengths U[150, 499]", ours_r, sdpa_r, useful, "TFLOP/s", "library = SDPA with the reference's dense key-padding mask (HF:758-762); FLOPs of valid frames only")
del qs, ks, vs, q4, k4, v4

# ---------------------------------------------------------------------------------------------- log_softmax / LayerNorm
rows_w, width = 128 * 499, 3184
logits = torch.randn(rows_w, width, device=dev)
out = torch.empty_like(logits)
am = torch.empty(rows_w, dtype=torch.int32, device=dev)
mx = torch.empty(rows_w, device=dev)
ours = timeit([lambda: ops.log_softmax_wide(logits, width, rows_w, width, out, width, am, mx)], iters=10)
lib_ms = timeit([lambda: torch.log_softmax(logits, -1)], iters=10)
row("log_softmax 63 872 x 3 184 fp32 (+argmax, max)", ours, lib_ms, 2.0 * rows_w * width * 4, "GB/s", "library = torch.log_softmax (argmax would be a second pass)")
del logits, out

xs = [torch.randn(M, 1024, device=dev) for _ in range(ROTATE)]
gamma, beta = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
o16 = [torch.empty(M, 1024, device=dev, dtype=torch.bfloat16) for _ in range(ROTATE)]
ours = timeit([lambda x=x, o=o: ops.layernorm_rows(x, M, 1024, 1024, gamma, beta, 1e-5, out_bf16=o, ld_bf16=1024) for x, o in zip(xs, o16)])
lib_ms = timeit([lambda x=x: torch.nn.functional.layer_norm(x, (1024,), gamma, beta, 1e-5).bfloat16() for x in xs])
row("LayerNorm 15 968 x 1 024 fp32 -> bf16", ours, lib_ms, M * 1024 * 6.0, "GB/s", "library = F.layer_norm + cast")
del xs, o16

# ---------------------------------------------------------------------------------------------- CTC (config 2 sizes: 8 and 64 utterances, 37 heads)
def synthetic_labels(frames, n_classes, seed, fraction=0.25):
    generator = torch.Generator().manual_seed(seed)
    label_lengths = (frames.double() * fraction).floor().long().clamp_min(1)
    labels = torch.zeros(len(frames), int(label_lengths.max()), dtype=torch.long)
    for index, length in enumerate(label_lengths.tolist()):
        labels[index, :length] = torch.randint(1, n_classes, (length,), generator=generator)
    return labels, label_lengths


for n_utt in (8, 64):
    frames = 749
    input_lengths = torch.randint(150, frames + 1, (n_utt,), generator=torch.Generator().manual_seed(4))
    input_lengths[0] = frames
    classes = [4] * 36 + [501]
    log_probs, labels_l, lens_l = [], [], []
    for head, c in enumerate(classes):
        log_probs.append(torch.log_softmax(torch.randn(frames, n_utt, c, device=dev), -1))
        labels, label_lengths = synthetic_labels(input_lengths, c, seed=head)
        labels_l.append(labels.to(dev))
        lens_l.append(label_lengths.to(dev))
    il = input_lengths.to(dev)
    problem = ops.CtcProblem(log_probs, labels_l, lens_l, il, batch_first=False, need_grad=True)
    scale = torch.ones(len(classes), device=dev)
    valid = int(input_lengths.sum())
    lp_bytes = float(sum(valid * c * 4 for c in classes))

    def ours_step():
        problem.forward()
        problem.backward(scale)

    leaves = [lp.detach().clone().requires_grad_(True) for lp in log_probs]

    def library_step():
        total = 0
        for lp, labels, label_lengths in zip(leaves, labels_l, lens_l):
            total = total + torch.nn.functional.ctc_loss(lp, labels, il, label_lengths, reduction="sum", zero_infinity=True)
        torch.autograd.grad(total, leaves)

    ours = timeit([ours_step], iters=6)
    lib_ms = timeit([library_step], iters=3, warm=1)
    row(f"CTC loss + gradient, 37 heads, {n_utt} utterances, T' <= 749", ours, lib_ms, 2.0 * lp_bytes, "GB/s", "library = 37 x F.ctc_loss + autograd; bytes = log-probs read + gradient written")

print(f"| kernel | ours (us) | library (us) | speed-up | ours | library | unit | note |")
print("|---|---:|---:|---:|---:|---:|---|---|")
for r in rows:
    print(f"| {r['name']} | {r['ours_us']:.1f} | {r['library_us']:.1f} | {r['speedup']:.2f}x | {r['ours_rate']:.1f} | {r['library_rate']:.1f} | {r['unit']} | {r['note']} |")
print()
print(f"measured peaks on this pool: HBM {peaks['hbm_gbs']:.0f} GB/s, bf16 {peaks['bf16_tflops']:.0f} (burst) / {peaks['bf16_tflops_sustained']:.0f} (sustained) TFLOP/s")
print(json.dumps(rows))
