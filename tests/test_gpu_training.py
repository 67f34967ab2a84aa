"""-m gpu: the training-side kernels and the end-to-end training step (forward + multi-head CTC + backward)
against torch autograd on the CPU / the CPU oracle, on identical seeded inputs and weights.

Tolerances.  fp32 kernels: 1e-3 relative (measured ~1e-6).  Kernels with bf16 operands (all tcgen05 GEMMs,
attention): each result is compared with an fp32 computation on the SAME bf16-rounded inputs, so only the
accumulation order and the bf16 rounding of intermediates differ: 2e-2 of the tensor's range.  End-to-end
gradients (bf16 activations and output gradients, fp32 accumulation and fp32 residual-gradient stream) are
compared per parameter tensor by ``||g - g_ref|| / ||g_ref||`` against the fp32 oracle: 5e-2.
"""
import math

import pytest
import torch
from torch.nn import functional as F

from oracle import restatement
from tests import helpers

pytestmark = pytest.mark.gpu

DEV = "cuda"
GRAD_TOL = 5e-2


def _ops():
    from allophant_b200 import ops

    return ops


def range_err(value, reference):
    reference = reference.double().cpu()
    return float((value.double().cpu() - reference).abs().max() / reference.abs().max().clamp_min(1e-12))


def norm_err(value, reference):
    reference = reference.double().cpu()
    return float((value.double().cpu() - reference).norm() / reference.norm().clamp_min(1e-20))


# ------------------------------------------------------------------------------------------ GEMM forms
@pytest.mark.parametrize("rows,k,n", [(1000, 1024, 4096), (333, 3072, 1024), (998, 152, 1088), (130, 64, 640), (77, 640, 64)])
def test_gemm_dgrad_form(rows, k, n):
    """dX = dY @ W with W read in its forward [out][in] layout as an MN-major B operand."""
    ops = _ops()
    torch.manual_seed(rows + k + n)
    dy = (torch.randn(rows, k, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(k, n, device=DEV) * 0.05).bfloat16()
    out = torch.empty(rows, n, device=DEV, dtype=torch.float32)
    ops.run_gemm(ops.make_dgrad_args(dy, w, rows=rows, ld_dy=k, k=k, n=n, ld_w=n, out_f32=out, ld_f32=n))
    reference = dy.float() @ w.float()
    assert range_err(out, reference) < 1e-4
    # accumulate onto an existing gradient + bf16 output with the GELU' epilogue
    pre = torch.randn(rows, n, device=DEV).bfloat16()
    out16 = torch.empty(rows, n, device=DEV, dtype=torch.bfloat16)
    ops.run_gemm(ops.make_dgrad_args(dy, w, rows=rows, ld_dy=k, k=k, n=n, ld_w=n, gelu_bwd=pre, ld_gelu_bwd=n, out_bf16=out16, ld_bf16=n))
    x = pre.float().requires_grad_(True)
    F.gelu(x).backward(reference)
    assert range_err(out16, x.grad) < 1e-2
    acc = torch.randn(rows, n, device=DEV)
    expected = acc + reference * 0.25
    ops.run_gemm(ops.make_dgrad_args(dy, w, rows=rows, ld_dy=k, k=k, n=n, ld_w=n, scale=0.25, resid=acc, ld_resid=n, out_f32=acc, ld_f32=n))
    assert range_err(acc, expected) < 1e-4


@pytest.mark.parametrize("rows,m,n", [(5992, 1024, 1024), (1000, 4096, 1024), (998, 1024, 4096), (333, 3072, 1024), (130, 152, 1088), (64, 64, 640), (4000, 8, 128)])
def test_gemm_wgrad_form(rows, m, n):
    """dW = dY^T X with both operands frame-major (MN-major A and B), any number of frames, split-K on long K."""
    ops = _ops()
    torch.manual_seed(rows + m + n)
    dy = (torch.randn(rows, m, device=DEV) * 0.5).bfloat16()
    x = (torch.randn(rows, n, device=DEV) * 0.5).bfloat16()
    out = torch.full((m, n), float("nan"), device=DEV, dtype=torch.float32)
    ops.run_gemm(ops.make_wgrad_args(dy, x, out, rows=rows, m=m, ld_dy=m, n=n, ld_x=n, ld_out=n, scale=0.5))
    reference = (dy.float().T @ x.float()) * 0.5
    assert range_err(out, reference) < 1e-4


def test_gemm_wgrad_strided_operands():
    """Operands that are column blocks of wider matrices (the heads' composition gradient)."""
    ops = _ops()
    torch.manual_seed(5)
    rows, m, n = 700, 64, 640
    dy_full = (torch.randn(rows, 128, device=DEV) * 0.5).bfloat16()
    x_full = (torch.randn(rows, 800, device=DEV) * 0.5).bfloat16()
    out = torch.empty(m, n, device=DEV, dtype=torch.float32)
    ops.run_gemm(ops.make_wgrad_args(dy_full, x_full[:, 160:], out, rows=rows, m=m, ld_dy=128, n=n, ld_x=800, ld_out=n))
    reference = dy_full[:, :m].float().T @ x_full[:, 160 : 160 + n].float()
    assert range_err(out, reference) < 1e-4


def test_forward_gemm_keeps_pre_activation():
    ops = _ops()
    torch.manual_seed(9)
    m, n, k = 777, 4096, 1024
    a = (torch.randn(m, k, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(n, device=DEV)
    act = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    ops.run_gemm(ops.make_gemm_args(a, w, a_rows=m, a_inner=k, a_row_stride=k, bias=bias, gelu=True, out_bf16=act, ld_bf16=n, aux_bf16=pre, ld_aux=n))
    reference = a.float() @ w.float().T + bias
    assert range_err(pre, reference) < 1e-2
    assert range_err(act, F.gelu(reference)) < 1e-2


# ------------------------------------------------------------------------------------------ elementwise / reductions
@pytest.mark.parametrize("cols,x_dtype,dy_dtype", [(1024, torch.float32, torch.float32), (1024, torch.float32, torch.bfloat16), (512, torch.bfloat16, torch.float32)])
def test_layernorm_backward(cols, x_dtype, dy_dtype):
    ops = _ops()
    torch.manual_seed(cols)
    rows = 1234
    x = (torch.randn(rows, cols, device=DEV) * 2 + 0.3).to(x_dtype)
    dy = torch.randn(rows, cols + 64, device=DEV).to(dy_dtype)  # strided gradient (a block of a wider matrix)
    gamma = torch.randn(cols, device=DEV)
    resid = torch.randn(rows, cols, device=DEV)
    dx = torch.empty(rows, cols, device=DEV)
    dgamma = torch.empty(cols, device=DEV)
    dbeta = torch.empty(cols, device=DEV)
    ops.layernorm_backward(x, cols, dy, cols + 64, rows, cols, gamma, 1e-5, resid, cols, dx, cols, dgamma, dbeta)
    xr = x.float().cpu().requires_grad_(True)
    gr = gamma.cpu().requires_grad_(True)
    br = torch.zeros(cols, requires_grad=True)
    F.layer_norm(xr, (cols,), gr, br, 1e-5).backward(dy[:, :cols].float().cpu())
    assert range_err(dx, xr.grad + resid.cpu()) < 1e-4
    assert range_err(dgamma, gr.grad) < 1e-4
    assert range_err(dbeta, br.grad) < 1e-4
    # in-place residual accumulation, no parameter gradients
    inplace = resid.clone()
    ops.layernorm_backward(x, cols, dy, cols + 64, rows, cols, gamma, 1e-5, inplace, cols, inplace, cols, None, None)
    assert range_err(inplace, xr.grad + resid.cpu()) < 1e-4


def test_small_training_kernels():
    ops = _ops()
    torch.manual_seed(3)
    rows, cols = 998, 1024
    x = torch.randn(rows, cols, device=DEV)
    x16 = x.bfloat16()
    assert range_err(ops.colsum_bf16(x16, rows, cols, cols), x16.float().sum(0)) < 1e-5
    assert range_err(ops.colsum_f32(x, rows, cols, cols), x.sum(0)) < 1e-5
    # ragged widths / leading dimensions (scalar tail path), few rows, a strided column block of a wider matrix
    for r, c, ld in [(5, 37, 40), (4896, 3072, 3072), (130, 501, 504), (1, 8, 8)]:
        wide = torch.randn(r, ld, device=DEV)
        assert range_err(ops.colsum_f32(wide, r, c, ld), wide[:, :c].sum(0)) < 1e-5
        wide16 = wide.bfloat16()
        assert range_err(ops.colsum_bf16(wide16, r, c, ld), wide16[:, :c].float().sum(0)) < 1e-5
    block = torch.randn(700, 2048, device=DEV)
    assert range_err(ops.colsum_f32(block[:, 1024:], 700, 1024, 2048), block[:, 1024:].sum(0)) < 1e-5
    lengths = torch.tensor([300, 499], device=DEV, dtype=torch.int32)
    masked = x.clone()
    ops.mask_rows(masked, cols, rows, cols, lengths, 499)
    expected = x.clone()
    expected[300:499] = 0
    assert torch.equal(masked, expected)
    other = torch.randn(rows, cols + 32, device=DEV)
    summed = x.clone()
    ops.add_2d(summed, cols, other[:, 32:], cols + 32, rows, cols)
    assert torch.equal(summed, x + other[:, 32:])
    pre = torch.randn(rows, cols, device=DEV).bfloat16()
    out = torch.empty(rows, cols, device=DEV, dtype=torch.bfloat16)
    ops.gelu_backward_bf16(x, cols, pre, cols, rows, cols, out, cols)
    p = pre.float().requires_grad_(True)
    F.gelu(p).backward(x)
    assert range_err(out, p.grad) < 1e-2


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("seq,lengths", [(249, [249, 100]), (499, [499, 300, 1]), (130, [129, 130]), (64, [64])])
def test_attention_backward(seq, lengths):
    """dQ/dK/dV of the varlen attention against torch autograd (fp32, same bf16-rounded q/k/v/dO)."""
    ops = _ops()
    torch.manual_seed(seq)
    n_utt, heads, d = len(lengths), 4, 64
    hidden = heads * d
    scale = 0.125 * 1.4426950408889634
    q_raw = torch.randn(n_utt, heads, seq, d, device=DEV)
    k = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16()
    v = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16()
    q_scaled = (q_raw * scale).bfloat16()
    frames = torch.tensor(lengths, device=DEV, dtype=torch.int32)
    ctx = torch.zeros(n_utt * seq, hidden, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(n_utt * heads * seq, device=DEV, dtype=torch.float32)
    ops.attention(q_scaled.contiguous(), k.contiguous(), v.contiguous(), ctx, frames, n_utt, heads, seq, lse)

    d_ctx = torch.randn(n_utt * seq, hidden, device=DEV)
    valid = torch.arange(seq, device=DEV)[None, :] < frames[:, None]  # [N, T]
    d_ctx = (d_ctx.view(n_utt, seq, hidden) * valid[..., None]).view(n_utt * seq, hidden).bfloat16()
    dqkv = torch.full((n_utt * seq, 3 * hidden), float("nan"), device=DEV, dtype=torch.bfloat16)
    delta = torch.empty(n_utt * heads * seq, device=DEV, dtype=torch.float32)
    ops.attention_backward(q_scaled.contiguous(), k.contiguous(), v.contiguous(), ctx, d_ctx, lse, delta, dqkv, frames, n_utt, heads, seq)

    # reference in fp32 on the CPU: q_unscaled = stored q / (0.125 log2 e) * 0.125-scaling folded like HF (q * head_dim^-0.5)
    qr = (q_scaled.float().cpu() / scale).requires_grad_(True)  # the "unscaled projection output" the kernel differentiates w.r.t.
    kr = k.float().cpu().requires_grad_(True)
    vr = v.float().cpu().requires_grad_(True)
    scores = (qr * 0.125) @ kr.transpose(-1, -2)
    key_mask = valid.cpu()[:, None, None, :]
    probs = torch.softmax(scores.masked_fill(~key_mask, float("-inf")), -1)
    out = (probs @ vr).permute(0, 2, 1, 3).reshape(n_utt * seq, hidden)
    # forward parity (valid rows), log-sum-exp parity
    valid_rows = valid.cpu().reshape(-1)
    assert range_err(ctx[valid_rows.to(DEV)], out[valid_rows].detach()) < 2e-2
    lse_ref = torch.logsumexp(scores.masked_fill(~key_mask, float("-inf")), -1) * 1.4426950408889634
    lse_ours = lse.view(n_utt, heads, seq).cpu()
    for b, length in enumerate(lengths):
        assert float((lse_ours[b, :, :length] - lse_ref[b, :, :length].detach()).abs().max()) < 2e-2
    out.backward(d_ctx.float().cpu())
    ours = dqkv.float().cpu().view(n_utt, seq, 3, heads, d).permute(2, 0, 3, 1, 4)  # [3, N, heads, T, d]
    assert torch.isfinite(ours).all()
    for part, reference in enumerate((qr.grad, kr.grad, vr.grad)):
        scale = float(reference.abs().max())  # one scale per tensor: a single-frame utterance has an exactly zero dQ/dK
        for b, length in enumerate(lengths):
            error = float((ours[part, b, :, :length] - reference[b, :, :length]).abs().max()) / scale
            assert error < 2e-2, (b, part, error)
            if length < seq:
                assert float(ours[part, b, :, length:].abs().max()) == 0.0, (b, part)


# ------------------------------------------------------------------------------------------ positional conv
@pytest.mark.parametrize("hidden", [1024, 768])
def test_posconv_backward(hidden):
    """Data and weight gradients of the weight-normed grouped positional conv (HF:326-368) + GELU + residual: 16 groups of 64
    channels (XLS-R) through the diagonal-block GEMM, 16 groups of 48 (wav2vec2-base) through block-diagonal super groups of 192
    channels (data gradient) and one full weight-gradient GEMM per tap."""
    ops = _ops()
    from allophant_b200 import _lib

    torch.manual_seed(17)
    n_utt, seq, groups, taps = 2, 200, 16, 128
    cg = hidden // groups
    weight_v = torch.randn(hidden, cg, taps, device=DEV) * 0.02
    weight_g = weight_v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt() * (1 + 0.1 * torch.randn(1, 1, taps, device=DEV))
    x16 = (torch.randn(n_utt, seq, hidden, device=DEV) * 0.5).bfloat16()
    dy16 = (torch.randn(n_utt, seq, hidden, device=DEV) * 0.5).bfloat16()

    # reference (fp32 CPU, torch autograd through weight norm and the grouped conv; SamePad drops the last frame)
    xr = x16.float().cpu().requires_grad_(True)
    vr = weight_v.cpu().requires_grad_(True)
    gr = weight_g.cpu().requires_grad_(True)
    w = gr * vr / vr.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    y = F.conv1d(xr.transpose(1, 2), w, None, padding=taps // 2, groups=groups)[:, :, :seq].transpose(1, 2)
    y.backward(dy16.float().cpu())

    # data gradient
    packed = ops.pack_posconv_weight_dgrad(weight_g, weight_v)
    span = 64
    if cg != 64:  # block-diagonal super groups, as PackedEncoder.pos_w_dgrad builds them
        import math

        span = cg * 64 // math.gcd(cg, 64)
        wide = torch.zeros(hidden, taps, span, device=DEV, dtype=torch.bfloat16)
        offsets = (torch.arange(hidden, device=DEV) % span) // cg * cg
        columns = offsets[:, None] + torch.arange(cg, device=DEV)[None, :]
        wide.scatter_(2, columns[:, None, :].expand(hidden, taps, cg), packed.view(hidden, taps, cg))
        packed = wide.view(hidden, taps * span)
    dx = torch.zeros(n_utt * seq, hidden, device=DEV)
    dgrad = ops.make_gemm_args(
        dy16, packed, a_rows=seq, a_inner=hidden, a_row_stride=hidden, batch=n_utt, a_batch_stride=seq * hidden, mode=_lib.APH_GEMM_TAPS,
        tap_pad=taps // 2 - 1, n=hidden, k=taps * span, resid=dx, ld_resid=hidden, out_f32=dx, ld_f32=hidden, out_batch_rows=seq,
    )  # fmt: skip
    dgrad.taps_span = span
    ops.run_gemm(dgrad)
    assert range_err(dx.view(n_utt, seq, hidden), xr.grad) < 2e-2

    # weight gradient
    if cg == 64:
        raw = torch.empty(taps, hidden, 256, device=DEV)
        args = ops.make_wgrad_args(dy16, x16, raw, rows=seq, m=hidden, ld_dy=hidden, n=hidden, ld_x=hidden, ld_out=256)
        args.mode, args.n_taps, args.tap_pad = _lib.APH_GEMM_DIAG_TAPS, taps, taps // 2
        args.k_batch, args.a_batch_stride, args.b_seg_stride = n_utt, seq * hidden, seq * hidden
        args.out_batch_rows = hidden
        ops.run_gemm(args)
        grad_g, grad_v = ops.posconv_weight_backward(raw, weight_g, weight_v)
    else:
        raw = torch.empty(taps, hidden, hidden, device=DEV)
        for tap in range(taps):
            args = ops.make_wgrad_args(dy16, x16, raw[tap], rows=seq, m=hidden, ld_dy=hidden, n=hidden, ld_x=hidden, ld_out=hidden)
            args.k_batch, args.a_batch_stride, args.b_seg_stride = n_utt, seq * hidden, seq * hidden
            args.b_k_shift = tap - taps // 2
            ops.run_gemm(args)
        grad_g, grad_v = ops.posconv_weight_backward(raw, weight_g, weight_v, block_width=hidden)
    assert range_err(grad_v, vr.grad) < 2e-2
    assert range_err(grad_g, gr.grad) < 2e-2


# ------------------------------------------------------------------------------------------ end to end
@pytest.fixture(scope="module", params=["multitask_2layer", "hierarchical_2layer", "allophones_2layer"])
def training_case(request):
    from allophant_b200.dataset_processing import Batch

    fixture = helpers.load_golden(f"training_{request.param}")
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    return dict(name=request.param, fixture=fixture, spec=spec, oracle=oracle, model=model, batch=batch, audio=audio, lengths=lengths)


def _training_step(model, batch, fixture):
    """estimator.py:708-738 on the CUDA path: forward, multi-head CTC, loss normalisation, backward."""
    from allophant_b200.loss_functions import multi_head_ctc_loss

    for parameter in model.parameters():
        parameter.grad = None
    predictions = model(batch)
    predictions.outputs.pop("phone", None)
    names = list(predictions.outputs)
    losses = multi_head_ctc_loss(
        [predictions.outputs[name] for name in names],
        [fixture["labels"][name].cuda() for name in names],
        predictions.lengths,
        [fixture["label_lengths"][name].cuda() for name in names],
    )
    normaliser = sum(int(fixture["label_lengths"][name].sum()) for name in names)
    loss = losses.sum() / normaliser
    loss.backward()
    return loss.detach(), dict(zip(names, losses.detach().cpu().tolist())), predictions


def test_training_step_matches_reference(training_case):
    """Loss and every parameter gradient of one training step against the unmodified reference (golden
    fingerprints) and against the CPU oracle's full gradients."""
    fixture, model, oracle = training_case["fixture"], training_case["model"], training_case["oracle"]
    model.eval()  # eval()-mode arithmetic (no dropout) like the golden vectors; parameters still require grad
    loss, per_head, predictions = _training_step(model, training_case["batch"], fixture)
    assert torch.equal(predictions.lengths.cpu(), fixture["frames"])
    assert abs(float(loss) - fixture["loss"]) <= 2e-2 * abs(fixture["loss"]), (float(loss), fixture["loss"])
    for name, value in fixture["per_head"].items():
        assert abs(per_head[name] - value) <= 2e-2 * max(1.0, abs(value)), name

    _, _, reference = oracle.training_step(
        training_case["audio"], training_case["lengths"], fixture["labels"], fixture["label_lengths"], fixture["language_ids"]
    )
    named = dict(model.named_parameters())
    frozen = set(fixture["frozen"])
    worst = {}
    for name, parameter in named.items():
        if name in frozen:
            assert parameter.grad is None or float(parameter.grad.abs().max()) == 0.0, f"{name} is frozen in the reference"
            continue
        assert parameter.grad is not None, f"no gradient for {name}"
        assert parameter.grad.shape == reference[name].shape
        assert torch.isfinite(parameter.grad).all(), name
        scale = float(reference[name].norm())
        if scale < 1e-7:  # k_proj.bias: the softmax is invariant to it, the exact gradient is 0 (reference: ~1e-9 of round-off)
            assert float(parameter.grad.norm()) < 1e-3
            continue
        worst[name] = norm_err(parameter.grad, reference[name])
        # the golden fingerprint comes from the UNMODIFIED reference
        summary = fixture["gradient_summaries"][name]
        assert abs(float(parameter.grad.double().norm()) - summary["norm"]) <= GRAD_TOL * summary["norm"], name
    ranked = sorted(worst.items(), key=lambda item: -item[1])
    print(f"{training_case['name']}: loss {float(loss):.5f} (reference {fixture['loss']:.5f}); worst gradient deviations: "
          + ", ".join(f"{k.split('._model.')[-1]}={v:.3e}" for k, v in ranked[:6]))  # fmt: skip
    assert ranked[0][1] < GRAD_TOL, ranked[:10]


def test_training_forward_equals_inference_forward(training_case):
    """The graph-building forward (activations kept, pre-activations written) computes the same logits as the
    inference forward."""
    model, batch = training_case["model"], training_case["batch"]
    model.eval()
    with torch.enable_grad():
        training = model(batch, predict=True)
    with torch.inference_mode():
        inference = model(batch, predict=True)
    # The inference launch list folds the encoder LayerNorms into the GEMMs around them (engine.EncoderPlan._build), the
    # training list keeps them as kernels: same arithmetic, different bf16 rounding points.  With the fold switched off the
    # two lists are the same kernels on the same inputs and the logits are bit-identical.
    for name, value in inference.outputs.items():
        assert helpers.rel_err(training.outputs[name].detach(), value) < 1e-2, name
    import os

    previous = os.environ.get("APH_FOLD_LN")
    os.environ["APH_FOLD_LN"] = "0"
    try:
        model.acoustic_model._plans.clear()
        with torch.inference_mode():
            unfolded = model(batch, predict=True)
    finally:
        if previous is None:
            os.environ.pop("APH_FOLD_LN", None)
        else:
            os.environ["APH_FOLD_LN"] = previous
        model.acoustic_model._plans.clear()
    for name, value in unfolded.outputs.items():
        assert torch.equal(training.outputs[name].detach(), value), name


def test_unfrozen_feature_extractor_trains(training_case):
    """freeze_feature_encoder = false (or after the reference's UnfreezeSchedule fired): the convolutional feature
    extractor gets gradients too — conv weights / biases and their LayerNorms, all seven layers — against the oracle with
    the extractor unfrozen.  Same tolerance as the rest of the step."""
    if training_case["name"] != "multitask_2layer":
        pytest.skip("one architecture is enough")
    fixture, model, oracle = training_case["fixture"], training_case["model"], training_case["oracle"]
    model.eval()
    extractor = [(name, parameter) for name, parameter in model.named_parameters() if ".feature_extractor." in name]
    assert len(extractor) == 28 and not any(parameter.requires_grad for _, parameter in extractor)
    try:
        for _, parameter in extractor:
            parameter.requires_grad = True
        loss, _, _ = _training_step(model, training_case["batch"], fixture)
        gradients = {name: parameter.grad.detach().clone() for name, parameter in model.named_parameters() if parameter.grad is not None}
    finally:
        for _, parameter in extractor:
            parameter.requires_grad = False
            parameter.grad = None
    reference_loss, _, reference = oracle.training_step(
        training_case["audio"], training_case["lengths"], fixture["labels"], fixture["label_lengths"], fixture["language_ids"], freeze_feature_encoder=False
    )
    oracle.trainable_parameters(True)  # freeze it again for the other tests
    assert abs(float(loss) - float(reference_loss)) <= 2e-2 * abs(float(reference_loss))
    worst = {}
    for name, _ in extractor:
        assert name in gradients, f"no gradient for {name}"
        assert torch.isfinite(gradients[name]).all(), name
        worst[name] = norm_err(gradients[name], reference[name])
    ranked = sorted(worst.items(), key=lambda item: -item[1])
    print("unfrozen feature extractor, worst gradient deviations: " + ", ".join(f"{k.split('feature_extractor.')[-1]}={v:.3e}" for k, v in ranked[:6]))
    assert ranked[0][1] < GRAD_TOL, ranked[:10]
    # the rest of the model still gets the gradients it got with the frozen extractor
    for name in ("_acoustic_model._model.encoder.layers.0.attention.out_proj.weight", "_acoustic_model._model.feature_projection.projection.weight"):
        assert norm_err(gradients[name], reference[name]) < GRAD_TOL, name


def test_frozen_encoder_trains_heads_only(training_case):
    fixture, model = training_case["fixture"], training_case["model"]
    if training_case["name"] != "hierarchical_2layer":
        pytest.skip("one architecture is enough")
    for parameter in model.acoustic_model.parameters():
        parameter.requires_grad = False
    try:
        _training_step(model, training_case["batch"], fixture)
        _, _, reference = training_case["oracle"].training_step(
            training_case["audio"], training_case["lengths"], fixture["labels"], fixture["label_lengths"], fixture["language_ids"]
        )
        for name, parameter in model.named_parameters():
            if name.startswith("_acoustic_model"):
                assert parameter.grad is None
            elif name in reference and float(reference[name].norm()) > 1e-7:
                assert norm_err(parameter.grad, reference[name]) < GRAD_TOL, name
    finally:
        frozen = set(fixture["frozen"])
        for name, parameter in model.named_parameters():
            parameter.requires_grad = name not in frozen


def test_backward_after_overwrite_raises(training_case):
    model, batch = training_case["model"], training_case["batch"]
    first = model(batch)
    model(batch)
    with pytest.raises(RuntimeError, match="overwritten"):
        next(iter(first.outputs.values())).sum().backward()


def test_optimizer_holds_parameters_that_are_unfrozen_later(training_case):
    """``optimizer_from_config`` hands the optimizer ALL parameters like the reference (``estimator.py:982``): the feature
    extractor is frozen at construction (``freeze_feature_encoder`` defaults to true) and unfrozen by ``UnfreezeSchedule.step``
    afterwards — it must then be updated (and the packed conv operands refreshed), and the param_groups must be the
    reference's (one group holding every parameter, so that a reference optimizer state_dict loads)."""
    if training_case["name"] != "multitask_2layer":
        pytest.skip("one architecture is enough")
    from allophant_b200 import optim
    from allophant_b200.config import Architecture, ProjectionConfig
    from allophant_b200.network.acoustic_model import UnfreezeSchedule

    fixture, model = training_case["fixture"], training_case["model"]
    model.eval()
    extractor = [(name, parameter) for name, parameter in model.named_parameters() if ".feature_extractor." in name]
    assert not any(parameter.requires_grad for _, parameter in extractor)
    before = {name: parameter.detach().clone() for name, parameter in model.named_parameters()}
    architecture = Architecture(1, ProjectionConfig([]), None, optimizer=dict(algorithm="adam", learning_rate=1e-4), lr_schedule=None)
    wrapper = optim.optimizer_from_config(architecture, model)
    assert [len(group["params"]) for group in wrapper.param_groups] == [len(list(model.parameters()))]
    schedule = UnfreezeSchedule(feature_extractor=1)
    try:
        _training_step(model, training_case["batch"], fixture)
        wrapper.step(clip_norm=1.0)
        assert all(torch.equal(parameter, before[name]) for name, parameter in extractor)  # still frozen: untouched
        schedule.step(model.acoustic_model)
        assert all(parameter.requires_grad for _, parameter in extractor)
        first_logits = None
        for _ in range(2):
            _training_step(model, training_case["batch"], fixture)
            wrapper.step(clip_norm=1.0)
        moved = [name for name, parameter in extractor if not torch.equal(parameter, before[name])]
        assert len(moved) == len(extractor), sorted(set(n for n, _ in extractor) - set(moved))[:4]
        # the forward pass sees the updated conv operands (the pack is refilled, not stale)
        with torch.inference_mode():
            first_logits = model(training_case["batch"], predict=True).outputs["stress"].clone()
        state = wrapper.state_dict()["optimizer"]
        assert len(state["param_groups"][0]["params"]) == len(before)
    finally:
        with torch.no_grad():
            for name, parameter in model.named_parameters():
                parameter.copy_(before[name])
                parameter.grad = None
        for _, parameter in extractor:
            parameter.requires_grad = False
        from allophant_b200 import engine

        engine.bump_weight_generation()
    with torch.inference_mode():
        restored = model(training_case["batch"], predict=True).outputs["stress"]
    assert first_logits is not None and not torch.equal(first_logits, restored)


# ------------------------------------------------------------------------------------------ optimiser step
def test_fused_adam_and_clipping_match_torch():
    """FusedAdam / clip_grad_norm_ (multi-tensor kernels) against torch.optim.Adam / nn.utils.clip_grad_norm_ on the same
    parameters and gradients: fp32 arithmetic, 1e-6 relative."""
    from allophant_b200 import optim

    torch.manual_seed(11)
    shapes = [(1024, 1024), (4096,), (3, 5, 7), (1,), (300, 1024)] + [(17,)] * 60  # more tensors than one launch packs
    ours = [torch.nn.Parameter(torch.randn(*shape, device=DEV)) for shape in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    fused = optim.FusedAdam(ours, lr=3e-3, betas=(0.9, 0.98), weight_decay=0.01)
    shadow = torch.zeros(1024, 1024, device=DEV, dtype=torch.bfloat16)
    fused.shadow[ours[0]] = shadow
    reference = torch.optim.Adam(theirs, lr=3e-3, betas=(0.9, 0.98), weight_decay=0.01)
    for step in range(4):
        grads = [torch.randn_like(p) * (10.0 if step == 1 else 0.01) for p in ours]
        for p, q, g in zip(ours, theirs, grads):
            p.grad = g.clone()
            q.grad = g.clone()
        if step % 2 == 0:  # in-place clipping, then a plain step
            norm = optim.clip_grad_norm_(ours, 1.5)
            norm_ref = torch.nn.utils.clip_grad_norm_(theirs, 1.5)
            assert abs(float(norm) - float(norm_ref)) <= 1e-5 * float(norm_ref)
            for p, q in zip(ours, theirs):
                assert range_err(p.grad, q.grad) < 1e-6
            fused.step()
        else:  # clipping folded into the Adam pass
            torch.nn.utils.clip_grad_norm_(theirs, 1.5)
            fused.step(clip_norm=1.5)
        reference.step()
        for index, (p, q) in enumerate(zip(ours, theirs)):
            assert range_err(p, q) < 1e-5, (step, index)
    assert torch.equal(shadow, ours[0].detach().bfloat16())
    state = fused.state_dict()
    assert set(state["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(state["state"][0]["step"]) == 4.0
    torch.optim.Adam(theirs, lr=3e-3, betas=(0.9, 0.98)).load_state_dict(state)  # interchangeable with torch's state layout


def test_training_loop_with_fused_optimizer_reduces_the_loss(training_case):
    """A few full steps (forward, multi-head CTC, backward, clip + Adam, warm-up schedule): the packed bf16 operands follow the
    updated fp32 parameters and the loss goes down."""
    from allophant_b200 import optim

    if training_case["name"] != "multitask_2layer":
        pytest.skip("one architecture is enough")
    fixture, model = training_case["fixture"], training_case["model"]
    saved = {name: value.detach().clone() for name, value in model.state_dict().items()}
    try:
        parameters = [p for p in model.parameters() if p.requires_grad]
        # peak rate 0.01 * 1024^-0.5 * 2^-0.5 = 2.2e-4: Adam moves every weight by about the rate per step
        optimizer = optim.adam_from_config(parameters, model.d_model, model=model, warmup_steps=2, constant_steps=100, factor=0.01)
        losses = []
        for _ in range(4):
            loss, _, _ = _training_step(model, training_case["batch"], fixture)
            losses.append(float(loss))
            optimizer.step(clip_norm=1.0)
        print("losses", losses, "lr", optimizer.current_learning_rate())
        assert all(torch.isfinite(torch.tensor(losses)))
        assert losses[-1] < losses[0]
        # the bf16 operands written by the Adam kernel are the rounded fp32 masters
        packed = model.acoustic_model._packed
        layer0 = model.acoustic_model.model.encoder.layers[0]
        assert torch.equal(packed.layers[0]["w1"], layer0.feed_forward.intermediate_dense.weight.detach().bfloat16())
        assert torch.equal(packed.layers[0]["wqkv"][1024:2048], layer0.attention.k_proj.weight.detach().bfloat16())
        assert torch.equal(packed.layers[0]["bqkv"][:1024], layer0.attention.q_proj.bias.detach())
    finally:
        model.load_state_dict(saved)


@pytest.mark.parametrize("momentum,weight_decay,clip", [(0.0, 0.0, None), (0.9, 1e-3, None), (0.9, 0.0, 0.5)])
def test_fused_sgd_matches_torch(momentum, weight_decay, clip):
    """FusedSGD (config.py:300-312: torch.optim.SGD(lr, momentum, weight_decay)) against torch on the same gradients, three
    steps, more tensors than one launch holds; clipping folded into the step like ``clip_grad_norm_`` + ``step``."""
    from allophant_b200 import optim

    torch.manual_seed(0)
    shapes = [(33,), (128, 64), (7, 5, 3)] * 20  # 60 tensors: two launches
    ours = [torch.nn.Parameter(torch.randn(shape, device=DEV)) for shape in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    fused = optim.FusedSGD(ours, lr=0.05, momentum=momentum, weight_decay=weight_decay)
    reference = torch.optim.SGD(theirs, lr=0.05, momentum=momentum, weight_decay=weight_decay)
    for step in range(3):
        for a, b in zip(ours, theirs):
            gradient = torch.randn_like(a) * (step + 1)
            a.grad, b.grad = gradient.clone(), gradient.clone()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(theirs, clip)
        fused.step(clip_norm=clip)
        reference.step()
        for a, b in zip(ours, theirs):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-6)
