"""Host feeding: samplers and collation against vectors produced by the unmodified reference
(``oracle/make_golden_batching.py`` -> ``tests/golden/batching.pt``)."""
import os

import pytest
import torch

from allophant_b200 import batching
from allophant_b200.dataset_processing import Batch, BatchType, LabeledBatch, RawLabeledBatch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "batching.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_batch_type_values_match_reference():
    assert [(member.name, member.value) for member in BatchType] == [("UNLABELED", 0), ("RAW", 1), ("INDEXED", 2)]


def test_max_frame_and_skip_samplers(golden):
    seen_empty_first = False
    for case in golden["samplers"]:
        sampler = batching.MaxFrameBatchSampler(case["order"], case["budget"], case["lengths"])
        batches = [batch for batch in sampler]
        assert batches == case["batches"]
        seen_empty_first |= bool(batches) and batches[0] == []
        assert [batch for batch in batching.SkipBatchSampler(sampler, 3)] == case["skipped"]
        for batch in batches:
            if len(batch) > 1:  # the frame budget holds for every batch of more than one utterance
                assert len(batch) * int(case["lengths"][batch].max()) <= case["budget"]
    assert seen_empty_first  # the reference's quirk (first utterance over budget) is covered


def _entries(golden, kind):
    names = ["phoneme", "syl", "son"]
    out = []
    for entry in golden["collate"]["entries"]:
        length, language = torch.tensor(entry["audio"].shape[0]), torch.tensor(entry["language"])
        if kind == BatchType.UNLABELED:
            out.append(Batch(entry["audio"], length, language))
        elif kind == BatchType.INDEXED:
            out.append(
                LabeledBatch(entry["audio"], length, language, entry["indices"], entry["label_lengths"], {n: i for i, n in enumerate(names)})
            )
        else:
            out.append(RawLabeledBatch(entry["audio"], length, language, list(entry["raw"]), [entry["utterance_id"]]))
    return out


@pytest.mark.parametrize("pinned", [False, True])
def test_collate_matches_reference(golden, pinned):
    expected = golden["collate"]
    unlabeled = batching.build_batch(BatchType.UNLABELED, pinned)(_entries(golden, BatchType.UNLABELED))
    assert type(unlabeled) is Batch
    assert torch.equal(unlabeled.audio_features, expected["unlabeled"]["audio"])
    assert torch.equal(unlabeled.lengths, expected["unlabeled"]["lengths"]) and unlabeled.lengths.dtype == torch.long
    assert torch.equal(unlabeled.language_ids, expected["unlabeled"]["languages"])

    indexed = batching.build_batch(BatchType.INDEXED, pinned)(_entries(golden, BatchType.INDEXED))
    assert torch.equal(indexed.audio_features, expected["indexed"]["audio"])
    assert indexed.label_length_indices == expected["indexed"]["label_length_indices"]
    assert len(indexed.attribute_indices) == len(expected["indexed"]["attribute_indices"])
    for mine, theirs in zip(indexed.attribute_indices, expected["indexed"]["attribute_indices"]):
        assert list(mine) == list(theirs)
        assert all(torch.equal(mine[name], theirs[name]) for name in theirs)
    for mine, theirs in zip(indexed.label_lengths, expected["indexed"]["label_lengths"]):
        assert torch.equal(mine, theirs)

    raw = batching.build_batch(BatchType.RAW, pinned)(_entries(golden, BatchType.RAW))
    assert torch.equal(raw.audio_features, expected["raw"]["audio"])
    assert raw.raw_labels == expected["raw"]["raw_labels"] and raw.utterance_ids == expected["raw"]["utterance_ids"]
    order = torch.argsort(raw.language_ids, stable=True)
    ordered = RawLabeledBatch(
        raw.audio_features[order],
        raw.lengths[order],
        raw.language_ids[order],
        [[labels[i] for i in order.tolist()] for labels in raw.raw_labels],
        [raw.utterance_ids[i] for i in order.tolist()],
    )
    parts = list(ordered.split_by_language())
    assert len(parts) == len(expected["raw"]["split"])
    for (language, part), (e_language, e_audio, e_lengths, e_labels, e_ids) in zip(parts, expected["raw"]["split"]):
        assert int(language) == e_language and torch.equal(part.audio_features, e_audio) and torch.equal(part.lengths, e_lengths)
        assert part.raw_labels == e_labels and part.utterance_ids == e_ids


def test_collate_audio_threads_and_edges():
    generator = torch.Generator().manual_seed(3)
    utterances = [torch.randn(int(n), generator=generator) for n in torch.randint(1, 60000, (24,), generator=generator)]
    expected = torch.nn.utils.rnn.pad_sequence(utterances, True)
    for threads in (1, 0, 5):
        assert torch.equal(batching.collate_audio(utterances, pinned=False, n_threads=threads), expected)
    # the staging ring: four consecutive batches do not alias each other
    first = batching.collate_audio(utterances[:3])
    kept = first.clone()
    for _ in range(3):
        batching.collate_audio([u + 1 for u in utterances[:3]])
    assert torch.equal(first, kept)
    # non-contiguous / non-fp32 sources, a single utterance, the empty batch
    odd = [torch.arange(10, dtype=torch.float64)[::2], torch.arange(3, dtype=torch.int16)]
    assert torch.equal(batching.collate_audio(odd, False), torch.nn.utils.rnn.pad_sequence([o.float() for o in odd], True))
    assert batching.collate_audio([torch.ones(5)], False).shape == (1, 5)
    assert batching.collate_audio([], False).shape == (0, 0)


def test_shard_for_rank_covers_the_batch():
    lengths = torch.tensor([100, 900, 500, 300, 700])
    audio = torch.nn.utils.rnn.pad_sequence([torch.full((int(n),), float(i)) for i, n in enumerate(lengths)], True)
    batch = Batch(audio, lengths, torch.arange(5))
    seen = []
    for rank in range(2):
        shard = batching.shard_for_rank(batch, rank, 2)
        assert shard.audio_features.shape[1] == int(shard.lengths.max())
        seen += shard.language_ids.tolist()
    assert sorted(seen) == [0, 1, 2, 3, 4]


def test_live_batches_do_not_alias_each_other():
    """Any number of collated batches may be alive at once (the reference's ``_training_batch_accumulation`` holds
    ``accumulation_factor`` host batches before ``.to(device)``): every batch owns its buffer."""
    import torch

    from allophant_b200.batching import build_batch
    from allophant_b200.dataset_processing import Batch, BatchType

    collate = build_batch(BatchType.UNLABELED)
    batches = []
    for value in range(7):
        entries = [Batch(torch.full((100 + 10 * i,), float(value + 1)), torch.tensor(100 + 10 * i), torch.tensor(0)) for i in range(3)]
        batches.append(collate(entries))
    pointers = {batch.audio_features.data_ptr() for batch in batches}
    assert len(pointers) == len(batches)
    for value, batch in enumerate(batches):
        assert batch.audio_features.shape == (3, 120)
        for row in range(3):
            assert bool((batch.audio_features[row, : 100 + 10 * row] == value + 1).all())
            assert bool((batch.audio_features[row, 100 + 10 * row :] == 0).all())
