"""TEST INFRASTRUCTURE — golden vectors for the from-scratch pre-LN transformer acoustic model, made by
the UNMODIFIED reference (``allophant/network/acoustic_model.py:34-69, 552-759``, ``frontend.py``, ``padding.py``) through
``oracle/reference_shim.py``.

For every case: the reference ``Allophant`` is built from a ``TransformerAcousticModelConfig`` under a fixed seed, its
``state_dict`` is frozen together with seeded input features, the hidden states of ``acoustic_model.forward`` and the
log-probabilities of ``model(batch, tfi, predict=True)`` -> ``tests/golden/transformer_<case>.pt``.
Usage: ``python -m oracle.make_golden_transformer``.
"""
from __future__ import annotations

import os
from typing import Any, Dict

import torch

from . import make_golden, reference_shim, restatement

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES: Dict[str, Dict[str, Any]] = {
    # Linear frontend (LayerNorm -> Linear -> LeakyReLU), GLU conv stack with reflect padding, GELU transformer
    "linear_glu": dict(
        feature_size=40,
        lengths=[118, 77, 30],
        acoustic=dict(
            transformer=dict(feedforward_neurons=128, heads=4, activation="gelu", num_layers=2, dropout_rate=0.1, positional_embeddings=True),
            frontend=dict(architecture="linear", neurons=64, input_dropout=0.1),
            sequential_frontend=[
                dict(type="glu1d", out_channels=96, kernel=3, stride=2),
                dict(type="layer_norm", affine=True),
                dict(type="dropout", rate=0.1),
                dict(type="glu1d", out_channels=256, kernel=5, stride=1),
            ],
            elementwise_affine=False,
        ),
        output_layers=[0],
    ),
    # Direct frontend straight into a ReLU transformer with affine LayerNorms and no positional embeddings
    "direct_relu": dict(
        feature_size=256,
        lengths=[64, 64, 9, 41],
        acoustic=dict(
            transformer=dict(feedforward_neurons=192, heads=4, activation="relu", num_layers=2, dropout_rate=0.0, positional_embeddings=False),
            frontend=dict(architecture="direct", input_dropout=0.0),
            sequential_frontend=None,
            elementwise_affine=True,
        ),
        output_layers=[],
    ),
}


def acoustic_config(cfg: Any, options: Dict[str, Any]) -> Any:
    frontend = dict(options["frontend"])
    architecture = frontend.pop("architecture")
    frontend_config = cfg.LinearFrontendConfig(**frontend) if architecture == "linear" else cfg.DirectFrontendConfig(**frontend)
    sequential = None
    if options["sequential_frontend"] is not None:
        kinds = {"glu1d": cfg.Glu1dConfig, "layer_norm": cfg.LayerNormConfig, "dropout": cfg.DropoutConfig}
        sequential = cfg.SequentialFrontendConfig(
            [kinds[layer["type"]](**{k: v for k, v in layer.items() if k != "type"}) for layer in options["sequential_frontend"]]
        )
    return cfg.TransformerAcousticModelConfig(cfg.TransformerConfig(**options["transformer"]), frontend_config, sequential, options["elementwise_affine"])


def run_case(name: str, case: Dict[str, Any]) -> Dict[str, Any]:
    spec_options = dict(n_train_phonemes=12, weight_seed=5)
    spec = restatement.multitask_spec(**spec_options)
    spec.embedding_size = 64
    if case["output_layers"]:  # one attribute classifier additionally reads an intermediate layer (OUTPUT_<i>)
        first = spec.classes[0]
        spec.classes[0] = restatement.ClassSpec(first.name, first.size, list(first.dependencies) + [f"OUTPUT_{i}" for i in case["output_layers"]])
    ref, model, _ = make_golden.reference_model(spec, lambda cfg: acoustic_config(cfg, case["acoustic"]), case["feature_size"])
    with torch.no_grad():  # make the deep-copied layers distinguishable and the affine LayerNorms non-trivial
        generator = torch.Generator().manual_seed(17)
        for parameter in model._acoustic_model.parameters():
            parameter.add_(0.05 * torch.randn(parameter.shape, generator=generator))
    lengths = torch.tensor(case["lengths"])
    generator = torch.Generator().manual_seed(3)
    features = torch.randn(len(lengths), case["feature_size"], int(lengths.max()), generator=generator)
    features = features * restatement.mask_sequence(lengths)[:, None, :]
    language_ids = torch.zeros(len(lengths), dtype=torch.long)
    table = torch.from_numpy(spec.feature_table).long()
    batch = ref.batching.Batch(features.clone(), lengths, language_ids)
    with torch.no_grad():
        hidden_states, frames = model._acoustic_model(ref.batching.Batch(features.clone(), lengths, language_ids))
        outputs = model(batch, table[:9], predict=True)
        log_probabilities = {key: model.log_probabilities(value) for key, value in outputs.outputs.items()}
    return dict(
        case=dict(case, spec=spec_options),
        state_dict={key: value.clone() for key, value in model.state_dict().items()},
        features=features,
        lengths=lengths,
        frames=frames,
        target_feature_indices=table[:9],
        hidden_states=[state.clone() for state in hidden_states],
        log_probabilities=log_probabilities,
        output_lengths=outputs.lengths,
    )


def main() -> None:
    reference_shim.install()
    for name, case in CASES.items():
        result = run_case(name, case)
        path = os.path.join(OUT_DIR, f"transformer_{name}.pt")
        torch.save(result, path)
        print(name, "frames", result["frames"].tolist(), "hidden", [tuple(h.shape) for h in result["hidden_states"]][:2],
              "heads", len(result["log_probabilities"]), os.path.getsize(path), "bytes")


def max_pool_reference_behaviour() -> Dict[str, Any]:
    """What the UNMODIFIED reference does with a ``max_pool`` layer (``frontend.py:258-259``): the sequential frontend runs, but
    reports MORE frames than it produced, and the transformer acoustic model behind it raises on the key-padding mask.  Frozen in
    ``tests/golden/max_pool_reference_behaviour.json`` as the justification for rejecting such configurations."""
    import importlib
    import json

    ref = reference_shim.reference_modules()
    cfg = ref.config
    frontend = importlib.import_module("allophant.network.frontend")
    acoustic = importlib.import_module("allophant.network.acoustic_model")
    dataset = importlib.import_module("allophant.dataset_processing")
    torch.manual_seed(0)
    lengths = torch.tensor([30, 21, 9])
    sequential = frontend.SequentialFrontend.from_config(cfg.SequentialFrontendConfig([cfg.MaxPoolingConfig(2)]), 8)
    pooled = sequential(dataset.Batch(torch.randn(3, 8, 30), lengths, torch.zeros(3)))
    options = dict(
        transformer=dict(feedforward_neurons=128, heads=4, activation="gelu", num_layers=1, dropout_rate=0.0, positional_embeddings=True),
        frontend=dict(architecture="linear", neurons=64, input_dropout=0.0),
        sequential_frontend=[dict(type="glu1d", out_channels=64, kernel=3, stride=1)],
        elementwise_affine=False,
    )
    config = acoustic_config(cfg, options)
    config.sequential_frontend.layers.append(cfg.MaxPoolingConfig(2))
    model = acoustic.TransformerAcousticModel.from_config(config, 40).eval()
    try:
        model(dataset.Batch(torch.randn(3, 40, 30), lengths, torch.zeros(3)))
        outcome = {"raised": None}
    except Exception as error:  # noqa: BLE001 - the point is to record whatever the reference does
        outcome = {"raised": type(error).__name__, "message": str(error)}
    record = {
        "input_frames": 30,
        "input_lengths": lengths.tolist(),
        "frontend_output_frames": int(pooled.audio_features.shape[-1]),
        "frontend_declared_lengths": pooled.lengths.tolist(),
        "transformer_model": outcome,
    }
    with open(os.path.join(OUT_DIR, "max_pool_reference_behaviour.json"), "w") as file:
        json.dump(record, file, indent=1)
    return record


if __name__ == "__main__":
    if os.environ.get("ONLY_MAX_POOL") != "1":
        main()
    print(max_pool_reference_behaviour())
