"""TEST INFRASTRUCTURE — golden vectors for the host-feeding row, made by the UNMODIFIED reference.

Runs ``allophant.batching.MaxFrameBatchSampler`` / ``SkipBatchSampler`` / ``_build_batch``
(``/root/reference/allophant/batching.py:94-215``) in this container through ``oracle/reference_shim.py`` on seeded
inputs and freezes the results in ``tests/golden/batching.pt``.  Usage: ``python -m oracle.make_golden_batching``.
"""
from __future__ import annotations

import os
import random

import torch

from . import reference_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "batching.pt")


def single_entries(reference_dp, seed: int, count: int, engines: int = 2):
    """Seeded single-utterance batches of the three kinds, as the reference's datasets emit them."""
    generator = torch.Generator().manual_seed(seed)
    rng = random.Random(seed)
    names = ["phoneme", "syl", "son"]
    entries = {"unlabeled": [], "indexed": [], "raw": []}
    plain = []
    for index in range(count):
        samples = int(torch.randint(1, 1500, (1,), generator=generator))
        audio = torch.randn(samples, generator=generator)
        language = rng.randrange(5)
        label_count = rng.randrange(1, 12)
        indices = [
            {name: torch.randint(1, 9, (label_count + shift,), generator=generator) for name in names} for shift in range(engines)
        ]
        label_lengths = [torch.tensor([label_count + shift] * len(names)) for shift in range(engines)]
        raw = [[[rng.choice("abcdefg") for _ in range(label_count + shift)]] for shift in range(engines)]
        plain.append(
            {
                "audio": audio,
                "language": language,
                "indices": indices,
                "label_lengths": label_lengths,
                "raw": raw,
                "utterance_id": f"utt{index}",
            }
        )
        length, language_id = torch.tensor(samples), torch.tensor(language)
        entries["unlabeled"].append(reference_dp.Batch(audio, length, language_id))
        entries["indexed"].append(
            reference_dp.LabeledBatch(audio, length, language_id, indices, label_lengths, {name: i for i, name in enumerate(names)})
        )
        entries["raw"].append(reference_dp.RawLabeledBatch(audio, length, language_id, [e for e in raw], [f"utt{index}"]))
    return plain, entries


def main() -> None:
    reference_shim.install()
    import allophant.batching as ref_batching
    import allophant.dataset_processing as ref_dp

    golden = {"samplers": [], "collate": {}}
    rng = random.Random(7)
    for case in range(6):
        count = [0, 1, 7, 40, 200, 64][case]
        lengths = torch.tensor([rng.randrange(1, 5000) for _ in range(count)], dtype=torch.long)
        order = list(range(count))
        rng.shuffle(order)
        budget = [100, 50, 6000, 12000, 30000, 5000][case]  # case 1: first utterance may exceed the budget -> empty batch
        batches = [batch for batch in ref_batching.MaxFrameBatchSampler(order, budget, lengths)]
        skipped = [batch for batch in ref_batching.SkipBatchSampler(ref_batching.MaxFrameBatchSampler(order, budget, lengths), 3)]
        golden["samplers"].append({"lengths": lengths, "order": order, "budget": budget, "batches": batches, "skipped": skipped})
    plain, entries = single_entries(ref_dp, 11, 9)
    golden["collate"]["entries"] = plain
    unlabeled = ref_batching._build_batch(ref_dp.BatchType.UNLABELED)(entries["unlabeled"])
    indexed = ref_batching._build_batch(ref_dp.BatchType.INDEXED)(entries["indexed"])
    raw = ref_batching._build_batch(ref_dp.BatchType.RAW)(entries["raw"])
    golden["collate"]["unlabeled"] = {"audio": unlabeled.audio_features, "lengths": unlabeled.lengths, "languages": unlabeled.language_ids}
    golden["collate"]["indexed"] = {
        "audio": indexed.audio_features,
        "lengths": indexed.lengths,
        "languages": indexed.language_ids,
        "attribute_indices": indexed.attribute_indices,
        "label_lengths": indexed.label_lengths,
        "label_length_indices": indexed.label_length_indices,
    }
    golden["collate"]["raw"] = {
        "audio": raw.audio_features,
        "raw_labels": raw.raw_labels,
        "utterance_ids": raw.utterance_ids,
        "split": [
            (int(language), part.audio_features, part.lengths, part.raw_labels, part.utterance_ids)
            for language, part in sorted_split(ref_dp, raw)
        ],
    }
    torch.save(golden, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


def sorted_split(reference_dp, raw):
    """``split_by_language`` of the batch re-ordered by language id (it splits CONSECUTIVE runs)."""
    order = torch.argsort(raw.language_ids, stable=True)
    ordered = reference_dp.RawLabeledBatch(
        raw.audio_features[order],
        raw.lengths[order],
        raw.language_ids[order],
        [[labels[i] for i in order.tolist()] for labels in raw.raw_labels],
        [raw.utterance_ids[i] for i in order.tolist()],
    )
    return list(ordered.split_by_language())


if __name__ == "__main__":
    main()
