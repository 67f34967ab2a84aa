"""TEST INFRASTRUCTURE — generates ``tests/golden/*.pt`` from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For every case it (1) builds the reference's own ``Allophant`` (``allophant/network/acoustic_model.py``)
through ``oracle/reference_shim.py`` under a fixed seed, (2) runs the reference's ``Estimator.predict``,
``CTCWrapper`` and ``GreedyCTCDecoder``, (3) rebuilds the same model with ``oracle/restatement.py``
and asserts the two agree (identical weights, outputs within 1e-5), and (4) stores the REFERENCE's
outputs.  The fixtures are what pins the restatement wherever ``/root/reference`` is absent.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_shim, restatement  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (spec kwargs, batch description).  Encoders are shrunk so a case runs in seconds on CPU;
# the full 24-layer shape is covered by ``xlsr300m_1x1s`` (outputs only: weights come from the seed).
CASES: Dict[str, Dict[str, Any]] = {
    "multitask_2layer": dict(
        spec=dict(n_train_phonemes=60, encoder_overrides=dict(num_hidden_layers=2)),
        lengths=[16000, 11111, 8000],
        inventory=25,
    ),
    "hierarchical_2layer": dict(
        spec=dict(n_train_phonemes=40, encoder_overrides=dict(num_hidden_layers=2), hierarchical=True),
        lengths=[12000, 9000],
        inventory=None,
    ),
    "allophones_2layer": dict(
        spec=dict(n_train_phonemes=30, encoder_overrides=dict(num_hidden_layers=2)),
        lengths=[10000, 10000, 7000],
        inventory=None,
        allophones=dict(n_languages=3, n_phones=50, seed=5),
    ),
    "xlsr300m_1x1s": dict(
        spec=dict(n_train_phonemes=60, encoder_overrides={}),
        lengths=[16000],
        inventory=25,
    ),
}


def synthetic_allophones(n_languages: int, n_phonemes: int, n_phones: int, seed: int) -> Dict[int, Dict[int, List[int]]]:
    """Per language: every phoneme of a random sub-inventory maps to 1-3 distinct shared phones."""
    rng = np.random.default_rng(seed)
    result: Dict[int, Dict[int, List[int]]] = {}
    for language in range(n_languages):
        inventory = sorted(rng.choice(n_phonemes, size=max(4, n_phonemes // 2), replace=False).tolist())
        mapping = {}
        for phoneme in inventory:
            count = int(rng.integers(1, 4))
            mapping[int(phoneme)] = sorted(int(p) for p in rng.choice(n_phones, size=count, replace=False))
        result[language] = mapping
    return result


def build_spec(case: Dict[str, Any]) -> restatement.OracleSpec:
    spec = restatement.multitask_spec(**case["spec"])
    allophones = case.get("allophones")
    if allophones is not None:
        n_phonemes = case["spec"]["n_train_phonemes"]
        spec.allophones = synthetic_allophones(allophones["n_languages"], n_phonemes, allophones["n_phones"], allophones["seed"])
        spec.n_phones = allophones["n_phones"]
        # with an allophone layer the composition table describes the shared PHONES
        rng = np.random.default_rng(77)
        table = rng.integers(0, 3, size=(spec.n_phones, len(restatement.PHOIBLE_FEATURES)))
        table[:3, :] = np.arange(3)[:, None]
        spec.feature_table = table
    return spec


def reference_model(spec: restatement.OracleSpec, acoustic_model_factory=None, feature_size: int = 1):
    """The reference's own Allophant over a synthetic indexer (a namespace with five fields suffices)."""
    import pandas as pd

    ref = reference_shim.reference_modules()
    reference_shim.set_encoder_overrides(**spec.encoder_overrides)
    cfg = ref.config
    features = restatement.PHOIBLE_FEATURES

    def attributes(table: np.ndarray, names: List[str]):
        frame = pd.DataFrame(
            {feature: [np.array([int(v)]) for v in table[:, i]] for i, feature in enumerate(features)},
            index=pd.Index(names, name="phoneme"),
        )
        categories = {feature: [str(v) for v in range(int(table[:, i].max()) + 1)] for i, feature in enumerate(features)}
        return ref.phonetic_features.ArticulatoryAttributes(frame, categories)

    table = np.asarray(spec.feature_table)
    phoneme_class = next(c for c in spec.classes if c.name == "phoneme")
    if spec.allophones is None:
        names = [f"p{i}" for i in range(table.shape[0])]
        full = attributes(table, names)
        indexer = types.SimpleNamespace(
            full_attributes=full,
            phonemes=pd.Index(names),
            composition_features=list(features),
            language_allophones=None,
            allophone_data=None,
        )
        phoneme_layer = cfg.PhonemeLayerType.SHARED
    else:
        phones = [f"ph{i}" for i in range(spec.n_phones)]
        shared = attributes(table, phones)
        mappings = ref.phonetic_features.LanguageAllophoneMappings(
            spec.allophones, [f"l{i}" for i in range(len(spec.allophones))], phones
        )
        indexer = types.SimpleNamespace(
            full_attributes=shared,
            phonemes=pd.Index([f"p{i}" for i in range(phoneme_class.size)]),
            composition_features=list(features),
            language_allophones=mappings,
            allophone_data=types.SimpleNamespace(shared_phone_indexer=shared),
        )
        phoneme_layer = cfg.PhonemeLayerType.ALLOPHONES

    entries = [cfg.ProjectionEntryConfig(c.name, list(c.dependencies)) for c in spec.classes]
    projection = cfg.ProjectionConfig(
        entries,
        phoneme_layer=phoneme_layer,
        acoustic_model_dropout=0.2,
        dependency_blanks=spec.dependency_blanks,
        embedding_composition=None if spec.embedding_size is None else cfg.EmbeddingCompositionConfig(spec.embedding_size),
    )
    architecture = types.SimpleNamespace(
        acoustic_model=cfg.Wav2Vec2PretrainedConfig("facebook/wav2vec2-xls-r-300m") if acoustic_model_factory is None else acoustic_model_factory(cfg),
        projection=projection,
        loss=cfg.CTCLossConfig(),
    )
    graph = ref.attribute_graph.AttributeGraph(
        ref.attribute_graph.AttributeNode(c.name, c.size, None, list(c.dependencies)) for c in spec.classes
    )
    torch.manual_seed(spec.weight_seed)
    model = ref.acoustic_model.Allophant.from_config(architecture, feature_size, 16000, graph, indexer, load_pretrained_weights=False)
    model.eval()
    estimator = types.SimpleNamespace(model=model)
    return ref, model, graph


def run_case(name: str, case: Dict[str, Any]) -> Dict[str, Any]:
    spec = build_spec(case)
    ref, model, graph = reference_model(spec)
    lengths = torch.tensor(case["lengths"], dtype=torch.long)
    n_utt, samples = len(lengths), int(lengths.max())
    audio = restatement.synthetic_audio(n_utt, samples, seed=0)
    audio = audio * restatement.mask_sequence(lengths)  # zero padding like batching.py:171-215 (pad_sequence)
    language_ids = torch.arange(n_utt) % max(1, len(spec.allophones or {0: 0}))
    inventory = case.get("inventory")
    tfi = None
    if inventory is not None:
        tfi = torch.from_numpy(np.random.default_rng(11).integers(0, 3, size=(inventory, len(restatement.PHOIBLE_FEATURES)))).long()

    batch = ref.batching.Batch(audio, lengths, language_ids)
    predict = ref.estimator.Estimator.predict  # unbound: only `self.model` is used (estimator.py:1035-1046)
    holder = types.SimpleNamespace(model=model)
    with torch.no_grad():
        predictions = predict.__wrapped__(holder, batch, tfi) if hasattr(predict, "__wrapped__") else predict(holder, batch, tfi)
        logits = model(batch, tfi, predict=True)
        hidden_states, frames = model.acoustic_model(batch)
        training_logits = model(batch, None, predict=False) if spec.allophones is not None else None

    # ---- the restatement must reproduce the reference bit-for-bit in weights and ~1e-6 in outputs
    oracle = restatement.OracleModel(spec)
    reference_state = model.state_dict()
    oracle_state = oracle.state_dict()
    assert sorted(reference_state) == sorted(oracle_state), (
        sorted(set(reference_state) ^ set(oracle_state))[:8],
        len(reference_state),
        len(oracle_state),
    )
    for key in reference_state:
        assert torch.equal(reference_state[key], oracle_state[key]), f"weights differ for {key}"
    oracle_outputs, oracle_frames = oracle.predict(audio, lengths, language_ids, tfi)
    assert torch.equal(oracle_frames, predictions.lengths)
    assert list(oracle_outputs) == list(predictions.outputs), (list(oracle_outputs)[:5], list(predictions.outputs)[:5])
    worst = 0.0
    for key, value in predictions.outputs.items():
        worst = max(worst, float((value - oracle_outputs[key]).abs().max()))
    assert worst < 1e-5, f"restatement deviates from the reference by {worst}"
    if training_logits is not None:
        oracle_hidden, _ = oracle.encode(audio, lengths)
        oracle_training = oracle.project(oracle_hidden, language_ids, None, predict=False)
        deviation = float((oracle_training["phoneme"] - training_logits.outputs["phoneme"]).abs().max())
        assert deviation < 1e-4, f"allophone mapping deviates by {deviation}"

    # ---- CTC loss and greedy decoding with the reference's own classes
    ctc = ref.loss_functions.CTCWrapper()
    decoder = ref.predictions.GreedyCTCDecoder()
    losses, labels_out, label_lengths_out, decoded = {}, {}, {}, {}
    for index, (key, value) in enumerate(logits.outputs.items()):
        if key == "phone":
            continue
        classes = value.shape[-1]
        labels, label_lengths = restatement.synthetic_labels(predictions.lengths, classes, seed=100 + index)
        losses[key] = ctc(value, labels, predictions.lengths, label_lengths)
        labels_out[key], label_lengths_out[key] = labels, label_lengths
    for key, value in predictions.outputs.items():
        hypotheses = decoder(value.transpose(1, 0).contiguous(), predictions.lengths)
        decoded[key] = [
            dict(tokens=h[0].tokens.clone(), timesteps=h[0].timesteps.clone(), score=float(h[0].score)) for h in hypotheses
        ]

    keep_heads = list(predictions.outputs)
    fixture = dict(
        case=name,
        case_config=case,
        checksum=restatement.state_checksum(reference_state),
        lengths=lengths,
        language_ids=language_ids,
        target_feature_indices=tfi,
        frames=predictions.lengths.clone(),
        head_order=keep_heads,
        log_probs={k: predictions.outputs[k].clone() for k in keep_heads},
        hidden_norms=[float(h.norm()) for h in hidden_states],
        last_hidden_slice=hidden_states[-1][:, :, :16].clone(),
        ctc_losses={k: float(v) for k, v in losses.items()},
        ctc_labels=labels_out,
        ctc_label_lengths=label_lengths_out,
        greedy=decoded,
        training_phoneme_logits=None if training_logits is None else training_logits.outputs["phoneme"].clone(),
        restatement_max_deviation=worst,
        versions=dict(torch=torch.__version__, transformers=__import__("transformers").__version__),
    )
    print(f"[{name}] heads={len(keep_heads)} frames={predictions.lengths.tolist()} restatement deviation={worst:.2e}")
    return fixture


def main() -> None:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    selected = sys.argv[1:] or list(CASES)
    for name in selected:
        fixture = run_case(name, CASES[name])
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(fixture, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
