"""The from-scratch pre-LN transformer acoustic model against vectors of the UNMODIFIED reference
(oracle/make_golden_transformer.py -> tests/golden/transformer_*.pt): module tree / state_dict / length arithmetic on
the CPU, hidden states and log-probabilities of the CUDA path on the GPU.

Tolerances (GPU): bf16 GEMM operands with fp32 accumulation and an fp32 residual stream — hidden states (LayerNorm
outputs, O(1)) within 3e-2 of their range on the valid frames, log-probabilities within 3e-2 of their range."""
import pytest
import torch

from tests import helpers

CASES = ["linear_glu", "direct_relu"]


@pytest.fixture(scope="module", params=CASES)
def golden(request):
    return helpers.load_golden(f"transformer_{request.param}")


def test_module_tree_matches_reference_state_dict(golden):
    model, _ = helpers.transformer_model_for_golden(golden, device="cpu")  # strict load inside
    acoustic = model._acoustic_model
    assert sorted(model.state_dict()) == sorted(golden["state_dict"])
    assert torch.equal(acoustic.downsampled_lengths(golden["lengths"]), golden["frames"])
    assert acoustic.d_model == 256 and acoustic.output_size == 256
    assert acoustic.hidden_state_count == len(golden["hidden_states"])
    from allophant_b200.config import Architecture  # noqa: F401  (config round trip)
    from allophant_b200 import config

    dumped = config._load_acoustic_model(
        dict(type="pre-ln-transformer", transformer=golden["case"]["acoustic"]["transformer"], frontend=golden["case"]["acoustic"]["frontend"])
    ).dump()
    assert config._load_acoustic_model(dumped).dump() == dumped
    with pytest.raises(RuntimeError, match="CUDA only"):
        from allophant_b200.dataset_processing import Batch

        acoustic(Batch(golden["features"], golden["lengths"], torch.zeros(len(golden["lengths"]))))


def test_same_seed_gives_the_reference_initialisation():
    """Modules are created in the reference's order, so a seed reproduces the reference's initial weights."""
    from allophant_b200.config import TransformerAcousticModelConfig
    from allophant_b200.network.acoustic_model import TransformerAcousticModel

    mapping = dict(
        type="pre-ln-transformer",
        transformer=dict(feedforward_neurons=256, heads=2, activation="gelu", num_layers=2),
        frontend=dict(architecture="linear", neurons=64),
        sequential_frontend={"layers": [dict(type="glu1d", out_channels=128, kernel=3, stride=2)]},
    )
    torch.manual_seed(11)
    ours = TransformerAcousticModel.from_config(TransformerAcousticModelConfig.load(mapping), 40)
    torch.manual_seed(11)
    linear = torch.nn.Linear(40, 64)  # frontend.py:171 is the first module with random parameters
    assert torch.equal(ours._frontend.linear.weight, linear.weight)
    layers = ours._transformer.layers
    assert torch.equal(layers[0].linear1.weight, layers[1].linear1.weight)  # nn.TransformerEncoder deep-copies ONE layer


def _range_err(value: torch.Tensor, reference: torch.Tensor) -> float:
    return float((value.double().cpu() - reference.double()).abs().max() / reference.double().abs().max())


@pytest.mark.gpu
def test_hidden_states_match_reference(golden):
    from allophant_b200.dataset_processing import Batch

    model, _ = helpers.transformer_model_for_golden(golden)
    batch = Batch(golden["features"].cuda(), golden["lengths"].cuda(), torch.zeros(len(golden["lengths"])).cuda())
    with torch.inference_mode():
        hidden_states, frames = model._acoustic_model(batch)
    assert torch.equal(frames.cpu(), golden["frames"])
    assert len(hidden_states) == len(golden["hidden_states"])
    for index, (ours, reference) in enumerate(zip(hidden_states, golden["hidden_states"])):
        assert ours.shape == reference.shape
        for utterance, length in enumerate(golden["frames"].tolist()):
            error = _range_err(ours[:length, utterance], reference[:length, utterance])
            assert error < 3e-2, (index, utterance, error)


@pytest.mark.gpu
def test_log_probabilities_match_reference(golden):
    from allophant_b200.dataset_processing import Batch
    from allophant_b200.estimator import Estimator

    model, _ = helpers.transformer_model_for_golden(golden)
    batch = Batch(golden["features"].cuda(), golden["lengths"].cuda(), torch.zeros(len(golden["lengths"])).cuda())
    estimator = Estimator.__new__(Estimator)
    with torch.inference_mode():
        predictions = model.predict_log_probabilities(batch, golden["target_feature_indices"].cuda())
    assert torch.equal(predictions.lengths.cpu(), golden["output_lengths"])
    assert sorted(predictions.outputs) == sorted(golden["log_probabilities"])
    del estimator
    for name, reference in golden["log_probabilities"].items():
        ours = predictions.outputs[name]
        assert ours.shape == reference.shape, name
        for utterance, length in enumerate(golden["frames"].tolist()):
            error = _range_err(ours[:length, utterance], reference[:length, utterance])
            assert error < 3e-2, (name, utterance, error)


@pytest.mark.gpu
def test_classifiers_train_on_a_frozen_transformer_and_unfrozen_raises(golden):
    from allophant_b200.dataset_processing import Batch

    model, _ = helpers.transformer_model_for_golden(golden)
    batch = Batch(golden["features"].cuda(), golden["lengths"].cuda(), torch.zeros(len(golden["lengths"]), dtype=torch.long).cuda())
    with pytest.raises(NotImplementedError, match="no backward pass"):
        model(batch)
    for parameter in model._acoustic_model.parameters():
        parameter.requires_grad = False
    outputs = model(batch)
    loss = sum(value.float().square().mean() for value in outputs.outputs.values())
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model._projection.parameters() if p.requires_grad)
