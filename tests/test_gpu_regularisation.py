"""-m gpu: train()-mode stochastic regularisation (dropout sites of the HF wav2vec2 encoder, LayerDrop, SpecAugment,
the dropout on the classifier inputs).

The CUDA kernels derive every keep mask from (seed, row, column) (``aph_common.cuh``: ``drop_hash``);
``tests/helpers.keep_mask`` restates that hash in numpy.  The kernels are therefore compared EXACTLY (same masks)
with fp32 torch computations, and the whole training step with the CPU oracle (the unmodified HF encoder in eval()
mode with the train()-mode ops applied through explicit masks, ``OracleModel.explicit_regularisation``).  Tolerances:
2e-2 of range for bf16-operand kernels (as in test_gpu_training.py); 1e-1 norm-relative for the end-to-end gradients.
The latter is read against the measured noise floor of the bf16 path on these tiny models: in eval() mode the worst
per-tensor deviation moves between 2.1e-2 and 5.6e-2 over eight audio seeds (profiles/r01_regularisation_noise_floor.log,
tests/diagnostics/gpu_noise_floor.py), and the train()-mode steps land in the same band (1.3e-2 ... 8.3e-2 over sites and seeds,
tests/diagnostics/gpu_regularisation_sites.py), while a wrong mask at any single site moves the gradients by >= 3e-1.
"""
import pytest
import torch

from oracle import restatement
from tests import helpers
from tests.test_gpu_training import _training_step, norm_err, range_err

pytestmark = pytest.mark.gpu

DEV = "cuda"
TRAIN_MODE_TOL = 1e-1


def _ops():
    from allophant_b200 import ops

    return ops


def test_dropout_kernels_use_the_restated_hash():
    ops = _ops()
    torch.manual_seed(0)
    rows, cols, ld = 333, 72, 80
    drop = ops.Dropout.site(0.3, 1234, 7)
    assert drop.threshold == round(0.3 * 65536) and abs(drop.scale - 65536 / (65536 - drop.threshold)) < 1e-6
    mask = helpers.keep_mask(drop, rows, cols)
    assert abs(float((mask == 0).float().mean()) - 0.3) < 0.02
    x = torch.randn(rows, ld, device=DEV)
    out = torch.full((rows, ld), 7.0, device=DEV)
    out16 = torch.zeros(rows, ld, device=DEV, dtype=torch.bfloat16)
    ops.dropout_2d(x, ld, rows, cols, drop, out_f32=out, ld_f32=ld, out_bf16=out16, ld_bf16=ld)
    expected = x[:, :cols].cpu() * mask
    assert torch.equal(out[:, :cols].cpu(), expected) and bool((out[:, cols:] == 7.0).all())
    assert torch.equal(out16[:, :cols].cpu(), expected.bfloat16())
    # SpecAugment rows replaced by the fill vector; in place
    row_mask = (torch.rand(rows, device=DEV) < 0.2).to(torch.uint8)
    fill = torch.randn(cols, device=DEV)
    y = x.clone()
    ops.dropout_2d(y, ld, rows, cols, drop, out_f32=y, ld_f32=ld, row_mask=row_mask, row_fill=fill)
    expected = torch.where(row_mask.bool().cpu()[:, None], fill.cpu()[None, :], x[:, :cols].cpu() * mask)
    assert torch.equal(y[:, :cols].cpu(), expected)
    # backward of the replaced rows
    d = torch.randn(rows, ld, device=DEV)
    d_fill = torch.full((cols,), float("nan"), device=DEV)
    d_before = d.clone()
    ops.masked_rows_backward(d, ld, rows, cols, row_mask, d_fill)
    assert range_err(d_fill, d_before[row_mask.bool(), :cols].sum(0)) < 1e-5
    assert bool((d[row_mask.bool(), :cols] == 0).all()) and torch.equal(d[~row_mask.bool()], d_before[~row_mask.bool()])
    # bf16 in place
    xb = torch.randn(rows, ld, device=DEV).bfloat16()
    before = xb.clone()
    ops.dropout_bf16_2d(xb, ld, rows, cols, drop)
    assert torch.equal(xb[:, :cols].cpu(), (before[:, :cols].float().cpu() * mask).bfloat16())
    assert torch.equal(xb[:, cols:], before[:, cols:])
    # different sites / seeds give different masks; threshold 0 is the identity
    assert not torch.equal(helpers.keep_mask(ops.Dropout.site(0.3, 1234, 8), rows, cols), mask)
    assert ops.Dropout.site(0.0, 1, 1) == ops.NO_DROPOUT


@pytest.fixture(params=[1, 2], ids=["blocks_of_64", "tile_pairs"])
def attention_kernel(request):
    from allophant_b200 import ops

    before = ops.set_attention_kernel(request.param)
    yield request.param
    ops.set_attention_kernel(before)


@pytest.mark.parametrize("rows,n,k", [(300, 1024, 1024), (517, 320, 4096)])
def test_gemm_epilogue_dropout(rows, n, k):
    """out = dropout(A W^T + b) + resid with the mask of (seed, row, col): both fp32-output epilogues."""
    ops = _ops()
    torch.manual_seed(rows)
    a = torch.randn(rows, k, device=DEV).bfloat16()
    w = (torch.randn(n, k, device=DEV) / k**0.5).bfloat16()
    bias = torch.randn(n, device=DEV)
    resid = torch.randn(rows, n, device=DEV)
    drop = ops.Dropout.site(0.1, 99, 3)
    out = torch.empty(rows, n, device=DEV)
    args = ops.make_gemm_args(a, w, a_rows=rows, a_inner=k, a_row_stride=k, bias=bias, resid=resid, ld_resid=n, out_f32=out, ld_f32=n)
    args.drop_threshold, args.drop_seed, args.drop_scale = drop.threshold, drop.seed, drop.scale
    ops.run_gemm(args)
    branch = a.float().cpu() @ w.float().cpu().T + bias.cpu()
    expected = branch * helpers.keep_mask(drop, rows, n) + resid.cpu()
    assert range_err(out, expected) < 2e-2
    dropped = helpers.keep_mask(drop, rows, n) == 0
    assert torch.equal(out.cpu()[dropped], resid.cpu()[dropped])  # dropped elements are EXACTLY the residual


@pytest.mark.parametrize("seq,lengths", [(200, [200, 77]), (131, [131, 1, 64])])
def test_attention_dropout_forward_and_backward(seq, lengths, attention_kernel):
    ops = _ops()
    torch.manual_seed(seq)
    n_utt, heads, d = len(lengths), 4, 64
    hidden = heads * d
    scale = 0.125 * 1.4426950408889634
    q_scaled = (torch.randn(n_utt, heads, seq, d, device=DEV) * scale).bfloat16().contiguous()
    k = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16().contiguous()
    v = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16().contiguous()
    frames = torch.tensor(lengths, device=DEV, dtype=torch.int32)
    drop = ops.Dropout.site(0.25, 5, 16)
    ctx = torch.zeros(n_utt * seq, hidden, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(n_utt * heads * seq, device=DEV)
    ops.attention(q_scaled, k, v, ctx, frames, n_utt, heads, seq, lse, drop)
    valid = torch.arange(seq, device=DEV)[None, :] < frames[:, None]
    d_ctx = (torch.randn(n_utt, seq, hidden, device=DEV) * valid[..., None]).view(n_utt * seq, hidden).bfloat16()
    dqkv = torch.full((n_utt * seq, 3 * hidden), float("nan"), device=DEV, dtype=torch.bfloat16)
    delta = torch.empty(n_utt * heads * seq, device=DEV)
    ops.attention_backward(q_scaled, k, v, ctx, d_ctx, lse, delta, dqkv, frames, n_utt, heads, seq, drop)

    qr = (q_scaled.float().cpu() / scale).requires_grad_(True)
    kr = k.float().cpu().requires_grad_(True)
    vr = v.float().cpu().requires_grad_(True)
    key_mask = valid.cpu()[:, None, None, :]
    probs = torch.softmax(((qr * 0.125) @ kr.transpose(-1, -2)).masked_fill(~key_mask, float("-inf")), -1)
    probs = probs * helpers.keep_mask(drop, n_utt * heads * seq, seq).view(n_utt, heads, seq, seq)
    out = (probs @ vr).permute(0, 2, 1, 3).reshape(n_utt * seq, hidden)
    valid_rows = valid.cpu().reshape(-1)
    assert range_err(ctx[valid_rows.to(DEV)], out[valid_rows].detach()) < 2e-2
    out.backward(d_ctx.float().cpu())
    ours = dqkv.float().cpu().view(n_utt, seq, 3, heads, d).permute(2, 0, 3, 1, 4)
    for part, reference in enumerate((qr.grad, kr.grad, vr.grad)):
        size = float(reference.abs().max())
        for b, length in enumerate(lengths):
            error = float((ours[part, b, :, :length] - reference[b, :, :length]).abs().max()) / size
            assert error < 2e-2, (b, part, error)


def test_spec_augment_mask_follows_the_hf_rules():
    """HF ``_compute_mask_indices``: spans = max(int(p * len / span + eps), min_masks) (clipped to what fits), each of
    ``span`` frames, starts distinct and inside the utterance; deterministic in the seed."""
    ops = _ops()
    seq, span, prob, min_masks = 499, 10, 0.075, 2
    frames = torch.tensor([499, 250, 60, 12, 9, 10], device=DEV, dtype=torch.int32)
    mask = torch.full((len(frames), seq), 9, device=DEV, dtype=torch.uint8)
    ops.spec_augment_mask(frames, seq, prob, span, min_masks, 4242, mask)
    again = torch.zeros_like(mask)
    ops.spec_augment_mask(frames, seq, prob, span, min_masks, 4242, again)
    other = torch.zeros_like(mask)
    ops.spec_augment_mask(frames, seq, prob, span, min_masks, 4243, other)
    assert torch.equal(mask, again) and not torch.equal(mask, other)
    host = mask.cpu()
    assert set(host.unique().tolist()) <= {0, 1}
    for row, length in zip(host, frames.tolist()):
        assert int(row[length:].sum()) == 0  # spans never leave the utterance
        low = max(int(prob * length / span), min_masks)
        high = max(int(prob * length / span + 1), min_masks)
        choices = length - (span - 1)
        low, high = min(low, max(choices, 0)), min(high, max(choices, 0))
        covered = int(row.sum())
        assert covered <= high * span
        if choices >= 1 and low >= 1:
            assert covered >= span  # at least one whole span
        if length == 9:
            assert covered == 0  # shorter than one span: nothing can be masked
        if length == 10:
            assert covered == 10  # exactly one start position
    with pytest.raises(ValueError):
        ops.spec_augment_mask(frames, 8, prob, span, min_masks, 1, mask)


@pytest.fixture(scope="module", params=["multitask_2layer", "hierarchical_2layer", "allophones_2layer"])
def case(request):
    from allophant_b200.dataset_processing import Batch

    fixture = helpers.load_golden(f"training_{request.param}")
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    return dict(name=request.param, fixture=fixture, oracle=oracle, model=model, batch=batch, audio=audio, lengths=lengths)


@pytest.mark.parametrize("skip_layers", [(False, False), (False, True)])
def test_train_mode_step_matches_oracle_with_the_same_masks(case, skip_layers):
    """One train()-mode step (all dropout sites, SpecAugment, LayerDrop, classifier-input dropout): loss and every
    parameter gradient against the oracle fed with the masks the kernels used."""
    fixture, model, oracle = case["fixture"], case["model"], case["oracle"]
    model.train()
    model._heads.skip_layers_override = list(skip_layers)
    try:
        torch.manual_seed(31)
        loss, per_head, predictions = _training_step(model, case["batch"], fixture)
        state = model._heads.last_regularisation
    finally:
        model._heads.skip_layers_override = None
        model.eval()
    stochastic, plan = state["stochastic"], state["plan"]
    assert stochastic is not None and plan.skipped == list(skip_layers) and plan.spec_active
    assert stochastic.hidden_dropout == 0.1 and stochastic.attention_dropout == 0.1 and stochastic.mask_time_prob == 0.075
    n_utt, seq = plan.n_utt, plan.seq
    cfg = plan.cfg
    masks = helpers.regularisation_masks(
        stochastic, n_utt, seq, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, plan.skipped, plan.spec_mask.cpu()
    )
    assert int(masks["spec"].sum()) > 0
    n_layers = cfg.num_hidden_layers
    blocks = {0: n_layers, **{column: index for index, column in model._heads.hidden_blocks.items()}}
    assert state["input_dropout"], "the test models are built with acoustic_model_dropout = 0.2"
    masks["classifier_input"] = {
        blocks[column]: helpers.keep_mask(drop, n_utt * seq, cfg.hidden_size).view(n_utt, seq, cfg.hidden_size)
        for column, drop in state["input_dropout"].items()
    }
    reference_loss, reference_heads, reference = oracle.training_step(
        case["audio"], case["lengths"], fixture["labels"], fixture["label_lengths"], fixture["language_ids"], regularisation=masks
    )
    assert abs(float(loss) - float(reference_loss)) <= 2e-2 * abs(float(reference_loss)), (float(loss), float(reference_loss))
    assert abs(float(reference_loss) - fixture["loss"]) > 1e-4 * abs(fixture["loss"])  # the masks did change the step
    for name, value in reference_heads.items():
        assert abs(per_head[name] - value) <= 2e-2 * max(1.0, abs(value)), name
    # "frozen" lists what got no gradient in the reference's eval()-mode step; masked_spec_embed is only used by SpecAugment
    frozen = set(fixture["frozen"]) - {"_acoustic_model._model.masked_spec_embed"}
    worst = {}
    for name, parameter in model.named_parameters():
        if name in frozen:
            continue
        skipped_layer = any(f".encoder.layers.{index}." in name for index, flag in enumerate(skip_layers) if flag)
        if skipped_layer:
            assert parameter.grad is not None and float(parameter.grad.abs().max()) == 0.0, name
            assert name not in reference or float(reference[name].abs().max()) == 0.0
            continue
        assert parameter.grad is not None, f"no gradient for {name}"
        assert name in reference, f"the oracle has no gradient for {name}"
        size = float(reference[name].norm())
        if size < 1e-7:
            assert float(parameter.grad.norm()) < 1e-3
            continue
        worst[name] = norm_err(parameter.grad, reference[name])
    ranked = sorted(worst.items(), key=lambda item: -item[1])
    print(f"{case['name']} skip={skip_layers}: loss {float(loss):.5f} (oracle {float(reference_loss):.5f}); worst: "
          + ", ".join(f"{k.split('._model.')[-1]}={v:.3e}" for k, v in ranked[:5]))  # fmt: skip
    assert "_acoustic_model._model.masked_spec_embed" in worst
    assert ranked[0][1] < TRAIN_MODE_TOL, ranked[:10]


def test_train_mode_is_seeded_and_eval_mode_is_untouched(case):
    fixture, model = case["fixture"], case["model"]
    model.eval()
    reference, _, _ = _training_step(model, case["batch"], fixture)
    model.train()
    try:
        torch.manual_seed(5)
        first, _, _ = _training_step(model, case["batch"], fixture)
        torch.manual_seed(5)
        second, _, _ = _training_step(model, case["batch"], fixture)
        torch.manual_seed(6)
        third, _, _ = _training_step(model, case["batch"], fixture)
    finally:
        model.eval()
    again, _, _ = _training_step(model, case["batch"], fixture)
    assert float(first) == float(second) and float(first) != float(third) and float(first) != float(reference)
    assert float(again) == float(reference)
    with torch.inference_mode():  # inference in train() mode stays deterministic (Estimator.predict switches to eval anyway)
        a = model(case["batch"], predict=True)
        b = model(case["batch"], predict=True)
    assert all(torch.equal(a.outputs[name], b.outputs[name]) for name in a.outputs)


def test_activation_dropout_matches_oracle_with_the_same_masks():
    """``activation_dropout > 0`` (HF ``intermediate_dropout`` behind the feed-forward GELU; 0 in the XLS-R configuration): the
    FFN1 epilogue drops the activation, the FFN2 data-gradient epilogue applies the same mask.  Only this site is enabled."""
    from allophant_b200.dataset_processing import Batch

    fixture = helpers.load_golden("training_multitask_2layer")
    case_config = dict(fixture["case_config"])
    case_config["spec"] = dict(case_config["spec"])
    overrides = dict(case_config["spec"].get("encoder_overrides") or {})
    overrides.update(activation_dropout=0.25, hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, layerdrop=0.0, mask_time_prob=0.0)
    case_config["spec"]["encoder_overrides"] = overrides
    spec = helpers.spec_for_case(case_config)
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    model._projection._acoustic_model_dropout.p = 0.0
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    model.train()
    try:
        torch.manual_seed(3)
        loss, _, _ = _training_step(model, batch, fixture)
        state = model._heads.last_regularisation
    finally:
        model.eval()
    stochastic, plan = state["stochastic"], state["plan"]
    assert stochastic.activation_dropout == 0.25 and stochastic.activation(0).threshold == 16384
    cfg = plan.cfg
    masks = helpers.regularisation_masks(
        stochastic, plan.n_utt, plan.seq, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, plan.skipped, None, cfg.intermediate_size
    )
    assert "activation.0" in masks and abs(float((masks["activation.1"] == 0).float().mean()) - 0.25) < 0.01
    reference_loss, _, reference = oracle.training_step(
        audio, lengths, fixture["labels"], fixture["label_lengths"], fixture["language_ids"], regularisation=masks
    )
    assert abs(float(loss) - float(reference_loss)) <= 2e-2 * abs(float(reference_loss))
    assert abs(float(reference_loss) - fixture["loss"]) > 1e-4 * abs(fixture["loss"])
    worst = {}
    for name, parameter in model.named_parameters():
        if parameter.grad is None or name not in reference or float(reference[name].norm()) < 1e-7:
            continue
        worst[name] = norm_err(parameter.grad, reference[name])
    ranked = sorted(worst.items(), key=lambda item: -item[1])
    print("activation dropout: worst " + ", ".join(f"{k.split('._model.')[-1]}={v:.3e}" for k, v in ranked[:4]))
    assert ranked[0][1] < TRAIN_MODE_TOL, ranked[:8]


def test_feature_axis_spec_augment_matches_oracle_with_the_same_masks():
    """``mask_feature_prob > 0`` (HF ``_mask_hidden_states``: spans along the hidden axis zeroed for every frame of an
    utterance, after the time mask): forward and gradients against the oracle fed with the masks the kernels drew."""
    from allophant_b200.dataset_processing import Batch

    fixture = helpers.load_golden("training_multitask_2layer")
    case_config = dict(fixture["case_config"])
    case_config["spec"] = dict(case_config["spec"])
    overrides = dict(case_config["spec"].get("encoder_overrides") or {})
    overrides.update(mask_feature_prob=0.1, mask_feature_length=10, hidden_dropout=0.0, attention_dropout=0.0, feat_proj_dropout=0.0, layerdrop=0.0)
    case_config["spec"]["encoder_overrides"] = overrides
    spec = helpers.spec_for_case(case_config)
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    model._projection._acoustic_model_dropout.p = 0.0
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    model.train()
    try:
        torch.manual_seed(9)
        loss, _, _ = _training_step(model, batch, fixture)
        state = model._heads.last_regularisation
    finally:
        model.eval()
    stochastic, plan = state["stochastic"], state["plan"]
    assert plan.feature_active and plan.spec_active and stochastic.mask_feature_prob == 0.1
    cfg = plan.cfg
    feature_mask = plan.feature_mask.cpu().bool()
    per_utterance = feature_mask.sum(1)
    assert feature_mask.shape == (plan.n_utt, cfg.hidden_size) and int(per_utterance.min()) >= 10 and int(per_utterance.max()) <= 110
    masks = helpers.regularisation_masks(
        stochastic, plan.n_utt, plan.seq, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, plan.skipped, plan.spec_mask.cpu()
    )
    masks["spec_feature"] = feature_mask
    reference_loss, _, reference = oracle.training_step(
        audio, lengths, fixture["labels"], fixture["label_lengths"], fixture["language_ids"], regularisation=masks
    )
    assert abs(float(loss) - float(reference_loss)) <= 2e-2 * abs(float(reference_loss))
    worst = {}
    for name, parameter in model.named_parameters():
        if parameter.grad is None or name not in reference or float(reference[name].norm()) < 1e-7:
            continue
        worst[name] = norm_err(parameter.grad, reference[name])
    ranked = sorted(worst.items(), key=lambda item: -item[1])
    print("feature-axis SpecAugment: worst " + ", ".join(f"{k.split('._model.')[-1]}={v:.3e}" for k, v in ranked[:4]))
    assert "_acoustic_model._model.masked_spec_embed" in worst
    assert ranked[0][1] < TRAIN_MODE_TOL, ranked[:8]


def test_post_ln_train_mode_matches_oracle_with_the_same_masks():
    """The post-LN encoder ordering in train() mode: the encoder-input dropout FOLLOWS the encoder LayerNorm there (HF
    ``Wav2Vec2Encoder``), every other site as in the stable ordering; one layer dropped by LayerDrop."""
    from allophant_b200.dataset_processing import Batch

    fixture = helpers.load_golden("training_multitask_2layer")
    case_config = dict(fixture["case_config"])
    case_config["spec"] = dict(case_config["spec"])
    overrides = dict(case_config["spec"].get("encoder_overrides") or {})
    overrides.update(do_stable_layer_norm=False)
    case_config["spec"]["encoder_overrides"] = overrides
    spec = helpers.spec_for_case(case_config)
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle)
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0) * restatement.mask_sequence(lengths)
    batch = Batch(audio.cuda(), lengths.cuda(), fixture["language_ids"].cuda())
    for skip_layers in ([False, False], [True, False]):
        model.train()
        model._heads.skip_layers_override = skip_layers
        try:
            torch.manual_seed(12)
            loss, _, _ = _training_step(model, batch, fixture)
            state = model._heads.last_regularisation
        finally:
            model._heads.skip_layers_override = None
            model.eval()
        stochastic, plan = state["stochastic"], state["plan"]
        cfg = plan.cfg
        assert not cfg.do_stable_layer_norm and plan.skipped == skip_layers
        masks = helpers.regularisation_masks(
            stochastic, plan.n_utt, plan.seq, cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, plan.skipped, plan.spec_mask.cpu()
        )
        blocks = {0: cfg.num_hidden_layers, **{column: index for index, column in model._heads.hidden_blocks.items()}}
        masks["classifier_input"] = {
            blocks[column]: helpers.keep_mask(drop, plan.n_utt * plan.seq, cfg.hidden_size).view(plan.n_utt, plan.seq, cfg.hidden_size)
            for column, drop in state["input_dropout"].items()
        }
        reference_loss, _, reference = oracle.training_step(
            audio, lengths, fixture["labels"], fixture["label_lengths"], fixture["language_ids"], regularisation=masks
        )
        assert abs(float(loss) - float(reference_loss)) <= 2e-2 * abs(float(reference_loss)), (float(loss), float(reference_loss))
        worst = {}
        for name, parameter in model.named_parameters():
            if parameter.grad is None or name not in reference or float(reference[name].norm()) < 1e-7:
                continue
            if any(f".encoder.layers.{index}." in name for index, flag in enumerate(skip_layers) if flag):
                assert float(parameter.grad.abs().max()) == 0.0, name
                continue
            worst[name] = norm_err(parameter.grad, reference[name])
        ranked = sorted(worst.items(), key=lambda item: -item[1])
        print(f"post-LN train() skip={skip_layers}: loss {float(loss):.5f} / {float(reference_loss):.5f}; worst "
              + ", ".join(f"{k.split('._model.')[-1]}={v:.3e}" for k, v in ranked[:4]))  # fmt: skip
        assert ranked[0][1] < TRAIN_MODE_TOL, ranked[:8]
