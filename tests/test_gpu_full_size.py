"""-m gpu: BASELINE.json's full-size configuration (configs[1]: XLS-R-300M shape, 24 layers, 32 x 10 s, 37 heads) through
size-independent properties — the CPU oracle needs minutes for this size, so instead of a value comparison the tests check
what must hold for ANY correct implementation of the path:

* every head's log-probabilities are normalised (logsumexp = 0) and finite on all valid frames; frame counts follow the
  convolution length formula (acoustic_model.py:832-835);
* utterances are independent: an utterance computed alone, inside a batch of different composition, or at a
  different batch position gives the same log-probabilities (bf16 tolerance: the GEMM tiling of a row does not depend on
  its neighbours, the attention masks other lengths) and the same greedy hypotheses wherever the argmax margin is clear;
* the path is deterministic: two runs are bit-identical;
* greedy decoding is the collapse of the per-frame argmax: tokens = unique-consecutive non-blank argmax, timesteps strictly
  increasing 1-based run starts, score = sum of the per-frame maxima (predictions.py:194-207).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def full_size():
    import bench
    from allophant_b200.dataset_processing import Batch

    estimator, tfi = bench.build_estimator(DEV)
    samples = bench.SECONDS * bench.SAMPLE_RATE
    generator = torch.Generator().manual_seed(7)
    audio = 0.1 * torch.randn(bench.BATCH, samples, generator=generator)
    lengths = torch.full((bench.BATCH,), samples, dtype=torch.long)
    lengths[3], lengths[17], lengths[31] = samples // 2, 16000, 400 + 3 * 320  # ragged: 5 s, 1 s and the shortest sensible clip
    audio = audio * (torch.arange(samples)[None, :] < lengths[:, None])
    batch = Batch(audio.to(DEV), lengths.to(DEV), torch.zeros(bench.BATCH, dtype=torch.long, device=DEV))
    with torch.inference_mode():
        predictions = estimator.predict(batch, tfi.to(DEV))
    return dict(estimator=estimator, tfi=tfi.to(DEV), audio=audio, lengths=lengths, batch=batch, predictions=predictions, Batch=Batch)


def test_log_probabilities_are_normalised_and_frames_follow_the_length_formula(full_size):
    predictions, lengths = full_size["predictions"], full_size["lengths"]
    frames = lengths.clone()
    for kernel, stride in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        frames = torch.div(frames - kernel, stride, rounding_mode="floor") + 1
    assert torch.equal(predictions.lengths.cpu(), frames)
    assert len(predictions.outputs) == 37 and int(frames.max()) == 499
    valid = (torch.arange(499)[:, None] < frames[None, :]).to(DEV)  # [T', N]
    for name, log_probs in predictions.outputs.items():
        assert log_probs.shape[:2] == (499, 32), name
        assert bool(torch.isfinite(log_probs[valid]).all()), name
        total = torch.logsumexp(log_probs.float(), -1)[valid]
        assert float(total.abs().max()) < 1e-4, (name, float(total.abs().max()))


def test_two_runs_are_bit_identical(full_size):
    with torch.inference_mode():
        again = full_size["estimator"].predict(full_size["batch"], full_size["tfi"])
    for name, value in full_size["predictions"].outputs.items():
        assert torch.equal(again.outputs[name], value), name


def test_utterances_are_independent_of_batch_composition(full_size):
    """Alone, re-ordered and in a smaller batch: same log-probabilities (<= 2e-2 of the head's range on valid frames)."""
    Batch, estimator, tfi = full_size["Batch"], full_size["estimator"], full_size["tfi"]
    audio, lengths, reference = full_size["audio"], full_size["lengths"], full_size["predictions"]
    picks = [3, 17, 31, 0]
    for group in ([3], [31, 0, 17], [17, 3]):
        longest = int(lengths[group].max())
        sub = Batch(audio[group, :longest].to(DEV), lengths[group].to(DEV), torch.zeros(len(group), dtype=torch.long, device=DEV))
        with torch.inference_mode():
            outputs = estimator.predict(sub, tfi)
        for position, index in enumerate(group):
            count = int(reference.lengths[index])
            assert int(outputs.lengths[position]) == count
            for name in ("phoneme", "syllabic", "continuant") if "syllabic" in reference.outputs else list(reference.outputs)[:3]:
                ours = outputs.outputs[name][:count, position].float()
                theirs = reference.outputs[name][:count, index].float()
                spread = float(theirs.max() - theirs.min())
                assert float((ours - theirs).abs().max()) <= 2e-2 * spread, (group, index, name)
    assert picks


def test_greedy_hypotheses_are_the_collapsed_argmax(full_size):
    from allophant_b200 import predictions as decoding

    reference = full_size["predictions"]
    hypotheses = decoding.decode_predictions(reference)
    assert sorted(hypotheses) == sorted(reference.outputs)
    for name in list(reference.outputs)[:4] + ["phoneme"]:
        log_probs = reference.outputs[name].float().transpose(0, 1).cpu()  # [N, T', C]
        for index in (0, 3, 17, 31):
            count = int(reference.lengths[index])
            best = log_probs[index, :count].max(-1)
            tokens, sizes = torch.unique_consecutive(best.indices, return_counts=True)
            starts = sizes.cumsum(0) - sizes + 1
            keep = tokens != 0
            [hypothesis] = hypotheses[name][index]
            assert hypothesis.tokens.tolist() == tokens[keep].tolist(), (name, index)
            assert hypothesis.timesteps.tolist() == starts[keep].tolist(), (name, index)
            assert abs(float(hypothesis.score) - float(best.values.sum())) <= 1e-3 * max(1.0, abs(float(best.values.sum())))
            assert hypothesis.timesteps.tolist() == sorted(set(hypothesis.timesteps.tolist()))


def test_full_size_gradients_of_shards_sum_to_the_full_batch_gradient():
    """BASELINE configs[2] at full model size (24 layers, allophone layer over 34 languages x 500 phones, 8 utterances of
    3-15 s): the UN-NORMALISED gradient of the summed CTC losses is additive over utterances, so the gradients of two
    half-batches must sum to the full-batch gradient — the property the data-parallel training step relies on (SURVEY.md §8e:
    sum all-reduce + one global normaliser).  eval() arithmetic; tolerance 2e-2 norm-relative per parameter tensor (bf16
    operands: the halves are padded to different lengths and tile differently)."""
    import bench
    from allophant_b200.dataset_processing import Batch
    from allophant_b200.loss_functions import multi_head_ctc_loss

    estimator, allophones = bench.build_training_estimator(DEV)
    model = estimator.model
    model.eval()
    generator = torch.Generator().manual_seed(3)
    count = bench.TRAIN_BATCH
    seconds = 3.0 + 12.0 * torch.rand(count, generator=generator)
    lengths = (seconds * bench.SAMPLE_RATE).long().sort(descending=True).values
    samples = int(lengths.max())
    audio = 0.1 * torch.randn(count, samples, generator=generator) * (torch.arange(samples)[None, :] < lengths[:, None])
    languages = torch.randint(0, bench.TRAIN_LANGUAGES, (count,), generator=generator)
    frames = model.downsampled_lengths(lengths)
    names = list(model.classes)
    labels, label_lengths = {}, {}
    for name in names:
        head_lengths = (frames.double() * 0.25).floor().long()
        head_labels = torch.zeros(count, int(head_lengths.max()), dtype=torch.long)
        for row, length in enumerate(head_lengths.tolist()):
            if name == "phoneme":
                inventory = torch.tensor(sorted(allophones[int(languages[row])]), dtype=torch.long) + 1
                head_labels[row, :length] = inventory[torch.randint(0, len(inventory), (length,), generator=generator)]
            else:
                head_labels[row, :length] = torch.randint(1, 4, (length,), generator=generator)
        labels[name], label_lengths[name] = head_labels, head_lengths
    parameters = {name: parameter for name, parameter in model.named_parameters() if parameter.requires_grad}

    def gradients(rows):
        rows = list(rows)
        longest = int(lengths[rows].max())
        batch = Batch(audio[rows, :longest].to(DEV), lengths[rows].to(DEV), languages[rows].to(DEV))
        for parameter in parameters.values():
            parameter.grad = None
        predictions = model(batch)
        predictions.outputs.pop("phone", None)
        order = list(predictions.outputs)
        losses = multi_head_ctc_loss(
            [predictions.outputs[name] for name in order],
            [labels[name][rows][:, : int(label_lengths[name][rows].max())].to(DEV) for name in order],
            predictions.lengths,
            [label_lengths[name][rows].to(DEV) for name in order],
        )
        total = losses.sum()
        total.backward()
        return float(total.detach()), {name: parameter.grad.detach().clone() for name, parameter in parameters.items() if parameter.grad is not None}

    full_loss, full = gradients(range(count))
    first_loss, first = gradients(range(0, count, 2))
    second_loss, second = gradients(range(1, count, 2))
    assert abs(first_loss + second_loss - full_loss) <= 2e-3 * abs(full_loss)
    assert len(full) > 400  # every trainable tensor of the 24-layer model got a gradient
    worst, worst_name = 0.0, ""
    for name, gradient in full.items():
        combined = first[name].double() + second[name].double()
        size = float(gradient.double().norm())
        if size < 1e-6 or name.endswith("attention.k_proj.bias"):
            # the key bias shifts every score of a row by the same amount and softmax does not see it: its gradient is analytically
            # zero, what the kernels produce is rounding noise (norm ~1e-6, sometimes just above the floor) with no additivity
            continue
        deviation = float((combined - gradient.double()).norm()) / size
        if deviation > worst:
            worst, worst_name = deviation, name
        assert torch.isfinite(gradient).all(), name
    print(f"full-size shard additivity: loss {full_loss:.2f} = {first_loss:.2f} + {second_loss:.2f}; worst gradient deviation {worst:.3e} ({worst_name})")
    assert worst < 2e-2, (worst, worst_name)
