"""TEST INFRASTRUCTURE — golden vectors for classifier heads with a multi-head-attention TIME LAYER
(``ProjectingMultiheadAttention``, ``allophant/network/acoustic_model.py:237-268``, built at 406-413), made by the UNMODIFIED
reference through ``oracle/reference_shim.py``.

The reference's ``Allophant`` (2-layer XLS-R-shaped encoder, seeded like the other golden cases) is built with time layers on
three attribute heads — one of them with sinusoidal positions, one with a single attention head — and a phoneme head that
depends on them; ``Estimator.predict`` runs on a ragged batch.  Frozen in ``tests/golden/time_layers_2layer.pt``: the projection's
``state_dict`` (the encoder's weights are the ones the plain ``multitask_2layer`` seed produces: the reference builds the encoder
first), the log-probabilities of every head and the frame counts.  Usage: ``python -m oracle.make_golden_time_layers``.
"""
from __future__ import annotations

import os
import types

import torch

from . import reference_shim, restatement

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "time_layers_2layer.pt")

TIME_LAYERS = {
    "stress": dict(num_heads=2, positional_embeddings=True),     # 3 categories + blank = 4 channels: two heads of 2
    "syllabic": dict(num_heads=1, positional_embeddings=False),  # one head of 4
    "nasal": dict(num_heads=4, positional_embeddings=False),     # four heads of 1
}
LENGTHS = [16000, 9000, 12345]
SPEC = dict(n_train_phonemes=20, encoder_overrides=dict(num_hidden_layers=2), weight_seed=2)


def main() -> None:
    ref = reference_shim.reference_modules()
    spec = restatement.multitask_spec(**SPEC)
    reference_shim.set_encoder_overrides(**spec.encoder_overrides)
    cfg = ref.config
    import numpy as np
    import pandas as pd

    features = restatement.PHOIBLE_FEATURES
    table = np.asarray(spec.feature_table)
    names = [f"p{i}" for i in range(table.shape[0])]
    frame = pd.DataFrame({feature: [np.array([int(v)]) for v in table[:, i]] for i, feature in enumerate(features)}, index=pd.Index(names, name="phoneme"))
    categories = {feature: [str(v) for v in range(int(table[:, i].max()) + 1)] for i, feature in enumerate(features)}
    full = ref.phonetic_features.ArticulatoryAttributes(frame, categories)
    indexer = types.SimpleNamespace(full_attributes=full, phonemes=pd.Index(names), composition_features=list(features), language_allophones=None, allophone_data=None)

    def time_layer(name: str):
        options = TIME_LAYERS.get(name)
        return None if options is None else cfg.MultiheadAttentionConfig(**options)

    # the phoneme head depends on the three time-layer heads (their probabilities are inputs of its projection)
    classes = [(c.name, c.size, list(c.dependencies) + (list(TIME_LAYERS) if c.name == "phoneme" else [])) for c in spec.classes]
    entries = [cfg.ProjectionEntryConfig(name, dependencies, time_layer(name)) for name, _, dependencies in classes]
    projection = cfg.ProjectionConfig(
        entries, phoneme_layer=cfg.PhonemeLayerType.SHARED, acoustic_model_dropout=0.2, dependency_blanks=spec.dependency_blanks,
        embedding_composition=cfg.EmbeddingCompositionConfig(spec.embedding_size),
    )  # fmt: skip
    architecture = types.SimpleNamespace(acoustic_model=cfg.Wav2Vec2PretrainedConfig("facebook/wav2vec2-xls-r-300m"), projection=projection, loss=cfg.CTCLossConfig())
    graph = ref.attribute_graph.AttributeGraph(ref.attribute_graph.AttributeNode(name, size, time_layer(name), dependencies) for name, size, dependencies in classes)
    torch.manual_seed(spec.weight_seed)
    model = ref.acoustic_model.Allophant.from_config(architecture, 1, 16000, graph, indexer, load_pretrained_weights=False)
    model.eval()

    lengths = torch.tensor(LENGTHS)
    audio = restatement.synthetic_audio(len(LENGTHS), max(LENGTHS), seed=0) * restatement.mask_sequence(lengths)
    batch = ref.batching.Batch(audio, lengths, torch.zeros(len(LENGTHS), dtype=torch.long))
    predict = ref.estimator.Estimator.predict
    holder = types.SimpleNamespace(model=model)
    with torch.no_grad():
        predictions = predict.__wrapped__(holder, batch, None) if hasattr(predict, "__wrapped__") else predict(holder, batch, None)

    # the encoder of this model is the encoder of the plain seeded model (it is constructed before the projection)
    plain = restatement.OracleModel(spec).state_dict()
    state = model.state_dict()
    for key, value in plain.items():
        if not key.startswith("_projection."):
            assert torch.equal(state[key], value), key
    result = dict(
        spec=SPEC,
        time_layers=TIME_LAYERS,
        classes=classes,
        lengths=lengths,
        frames=predictions.lengths,
        head_order=list(predictions.outputs),
        log_probs={name: value.clone() for name, value in predictions.outputs.items()},
        projection_state={key: value.clone() for key, value in state.items() if key.startswith("_projection.")},
    )
    torch.save(result, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(result["projection_state"]), "projection tensors;", {n: tuple(v.shape) for n, v in list(result["log_probs"].items())[:2]})


if __name__ == "__main__":
    main()
