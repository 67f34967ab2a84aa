"""CPU: host-side mirror of the reference interface (graph order, config schema, batch containers, layouts)."""
import random

import pytest
import torch

from oracle import restatement
from tests import helpers


def test_attribute_graph_order_matches_reference_order():
    from allophant_b200.attribute_graph import AttributeGraph, AttributeNode, DependencyCycleError

    # golden head orders come from the reference's own AttributeGraph.sort()
    for case in ("multitask_2layer", "hierarchical_2layer"):
        fixture = helpers.load_golden(case)
        spec = helpers.spec_for_case(fixture["case_config"])
        graph = AttributeGraph(AttributeNode(c.name, c.size, None, list(c.dependencies)) for c in spec.classes)
        assert [node.name for node in graph.sort()] == fixture["head_order"]
    # random DAGs: same order as the restated Tarjan post-order
    generator = random.Random(3)
    for _ in range(50):
        count = generator.randint(1, 12)
        classes = []
        for index in range(count):
            dependencies = ["OUTPUT"] + [f"n{j}" for j in range(index + 1, count) if generator.random() < 0.3]
            generator.shuffle(dependencies)
            classes.append(restatement.ClassSpec(f"n{index}", generator.randint(1, 5), dependencies))
        graph = AttributeGraph(AttributeNode(c.name, c.size, None, list(c.dependencies)) for c in classes)
        assert [n.name for n in graph.sort()] == [c.name for c in restatement.topological_order(classes)]
        restored = AttributeGraph.from_state(graph.state())
        assert [n.name for n in restored.sort()] == [n.name for n in graph.sort()]
    with pytest.raises(DependencyCycleError):
        list(AttributeGraph([AttributeNode("a", 1, None, ["b"]), AttributeNode("b", 1, None, ["a"])]).sort())


def test_config_schema_roundtrip_and_defaults():
    from allophant_b200.config import Config, PhonemeLayerType, ProjectionEntryConfig

    config = Config.default()
    assert len(config.nn.projection.classes) == 37
    assert config.nn.projection.classes[-1].name == ProjectionEntryConfig.PHONEME_LAYER
    assert config.nn.projection.embedding_composition.embedding_size == 640
    assert config.nn.projection.phoneme_layer == PhonemeLayerType.ALLOPHONES
    assert config.nn.projection.acoustic_model_dropout == 0.2
    assert config.nn.acoustic_model.model_id == "facebook/wav2vec2-xls-r-300m" and config.nn.acoustic_model.freeze_feature_encoder
    assert config.nn.loss.BLANK_OFFSET == 1
    dumped = config.dump()
    assert Config.load(dumped).dump() == dumped
    assert dumped["nn"]["acoustic_model"]["type"] == "wav2vec2-pretrained" and dumped["nn"]["loss"]["type"] == "CTC"
    minimal = {"nn": {"batch_size": 8, "projection": {"classes": [{"name": "phoneme"}]},
                      "acoustic_model": {"type": "wav2vec2-pretrained", "model_id": "facebook/wav2vec2-xls-r-300m"}}}  # fmt: skip
    loaded = Config.load(minimal)
    assert loaded.nn.projection.classes[0].dependencies == ["OUTPUT"] and loaded.nn.projection.dependency_blanks
    assert ProjectionEntryConfig.OUTPUT_PATTERN.match("OUTPUT_12").group(1) == "12"


def test_batch_containers_and_length_helpers():
    from allophant_b200.dataset_processing import Batch, LabeledBatch
    from allophant_b200.network.frontend import conv_length
    from allophant_b200.utils import mask_sequence

    lengths = torch.tensor([5, 3, 0])
    assert torch.equal(mask_sequence(lengths), restatement.mask_sequence(lengths))
    assert torch.equal(mask_sequence(lengths, inverse=True), restatement.mask_sequence(lengths, inverse=True))
    assert mask_sequence(lengths, 7, batch_first=False).shape == (7, 3)
    samples = torch.tensor([16000, 400, 399, 160000, 10])
    folded = samples
    for kernel, stride in zip(restatement.XLSR_300M["conv_kernel"], restatement.XLSR_300M["conv_stride"]):
        folded = conv_length(kernel, stride, use_padding=False)(folded)
    assert torch.equal(folded, restatement.conv_lengths(samples, restatement.XLSR_300M["conv_kernel"], restatement.XLSR_300M["conv_stride"]))
    batch = Batch(torch.zeros(2, 8), torch.tensor([8, 4]), torch.zeros(2, dtype=torch.long))
    assert len(batch) == 2 and batch.size() == 2 and "Features" in repr(batch)
    moved = batch.to("cpu", copy=True)
    assert isinstance(moved, Batch) and moved.audio_features.data_ptr() != batch.audio_features.data_ptr()
    labeled = LabeledBatch(batch.audio_features, batch.lengths, batch.language_ids, [{"phoneme": torch.ones(2, 3, dtype=torch.long)}],
                           [torch.tensor([[3, 2]])], {"phoneme": 0})  # fmt: skip
    moved = labeled.to("cpu")
    assert isinstance(moved, LabeledBatch) and moved.label_length_indices == {"phoneme": 0}
    assert [f.name for f in __import__("dataclasses").fields(LabeledBatch)] == [
        "audio_features", "lengths", "language_ids", "attribute_indices", "label_lengths", "label_length_indices"]  # fmt: skip


@pytest.mark.parametrize("case", ["multitask_2layer", "hierarchical_2layer", "allophones_2layer"])
def test_state_dict_layout_is_the_reference_layout(case):
    fixture = helpers.load_golden(case)
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle, device="cpu")  # construction + load_state_dict only
    ours = model.state_dict()
    reference = oracle.state_dict()
    assert sorted(ours) == sorted(reference)
    for key, value in reference.items():
        assert ours[key].shape == value.shape, key
        assert torch.equal(ours[key], value), key
    assert model.classes == [c.name for c in spec.classes]
    assert model.d_model == 1024 and model.feature_size == 1 and model.upscale_factor == 1
    assert model.l2_penalty() is None
    assert torch.equal(model.downsampled_lengths(torch.tensor([16000, 8000])), torch.tensor([49, 24]))
    # frozen feature encoder (default) and trainable rest
    assert not any(p.requires_grad for p in model.acoustic_model.model.feature_extractor.parameters())
    assert all(p.requires_grad for p in model.acoustic_model.model.encoder.parameters())
    # the forward path refuses to run off-GPU instead of falling back
    from allophant_b200.dataset_processing import Batch

    with pytest.raises(RuntimeError, match="CUDA"), torch.inference_mode():
        model(Batch(torch.zeros(1, 4000), torch.tensor([4000]), torch.zeros(1, dtype=torch.long)), predict=True)


def test_heads_layout_for_dependency_graph():
    fixture = helpers.load_golden("hierarchical_2layer")
    spec = helpers.spec_for_case(fixture["case_config"])
    oracle = restatement.OracleModel(spec)
    model, _ = helpers.cuda_model_for_spec(spec, oracle, device="cpu")
    runtime = model._heads
    runtime._build_layout()
    # X = [OUTPUT(1024) | 36 x softmax(4) | pad] -> 1168 columns rounded up to a multiple of 64
    assert runtime.ldx == 1216 and len(runtime.levels) == 2
    assert runtime.dep_cols["stress"] == (1024, 4) and len(runtime.dep_cols) == 36
    assert [s.name for s in runtime.levels[1].specs] == ["phoneme"] and runtime.levels[0].n_pad == 144
    assert runtime.levels[0].feeds_later == [s.name for s in runtime.levels[0].specs]


def test_error_behaviour_mirrors_reference():
    from allophant_b200.attribute_graph import AttributeGraph, AttributeNode
    from allophant_b200.config import EmbeddingCompositionConfig
    from allophant_b200.network.acoustic_model import HierarchicalProjection

    with pytest.raises(ValueError, match="reserved keyword"):
        HierarchicalProjection(1024, AttributeGraph([AttributeNode("OUTPUT", 3, None, ["OUTPUT"])]), 1)
    with pytest.raises(ValueError, match="requires a dependency"):
        HierarchicalProjection(1024, AttributeGraph([AttributeNode("a", 3, None, [])]), 1)
    with pytest.raises(ValueError, match="requires an attribute indexer"):
        HierarchicalProjection(1024, AttributeGraph([AttributeNode("phoneme", 3, None, ["OUTPUT"])]), 1,
                               embedding_composition_config=EmbeddingCompositionConfig(64))  # fmt: skip


def test_warmup_schedule_matches_the_reference():
    """WarmupScheduler (config.py:107-173) against learning rates produced by the unmodified reference class
    (tests/golden/warmup_schedule.json, generated through oracle/reference_shim.py)."""
    import json
    import os

    import torch

    from allophant_b200.optim import WarmupInfo, WarmupScheduler

    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "warmup_schedule.json")))
    optimizer = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    scheduler = WarmupScheduler(optimizer, WarmupInfo(golden["model_size"]), golden["warmup_steps"], golden["constant_steps"], golden["factor"])
    wanted = dict(zip(golden["steps"], golden["rates"]))
    step = 1
    assert scheduler.last_lr == wanted[1] and optimizer.param_groups[0]["lr"] == wanted[1]
    while step < max(wanted):
        scheduler.step()
        step += 1
        if step in wanted:
            assert scheduler.last_lr == wanted[step], step
            assert optimizer.param_groups[0]["lr"] == wanted[step]
    state = scheduler.state_dict()
    other = WarmupScheduler(optimizer, WarmupInfo(golden["model_size"]), golden["warmup_steps"], golden["constant_steps"], golden["factor"])
    other.load_state_dict(state)
    other.step()
    scheduler.step()
    assert other.last_lr == scheduler.last_lr


def test_prediction_files_round_trip(tmp_path):
    """JSON-lines prediction files, format 1.1.0 (allophant/predictions.py:29-186): metadata line + one record per utterance,
    plain and gzip, exclusive creation, edit operations with integer actions."""
    from allophant_b200 import predictions
    from allophant_b200.config import FeatureSet
    from allophant_b200.phonemes import Action

    metadata = predictions.PredictionMetaData("--beam 1", "common-voice", ["de", "es"], FeatureSet.PHOIBLE, {"phonemes": ["a", "b"]}, ["phoneme", "syl"], {"phoneme": ["a", "b"]})
    records = [
        predictions.UtterancePrediction("de", "utt0", {"phoneme": [["a", "b"]], "syl": [["+", "-"]]}, [["a", "b"]]),
        predictions.UtterancePrediction("es", "utt1", {"phoneme": [["t͡ʃ"]], "syl": [[]]}),
    ]
    for name in ("predictions.jsonl", "predictions.jsonl.gz"):
        path = tmp_path / name
        with predictions.JsonlWriter(path, metadata, gzip=None) as writer:
            for record in records:
                writer.write(record)
        with pytest.raises(FileExistsError):
            predictions.JsonlWriter(path, metadata, gzip=None).__enter__()
        with predictions.PredictionReader(path) as reader:
            assert reader.metadata == metadata and reader.metadata.format_version == (1, 1, 0)
            assert list(reader) == records
    import gzip
    import json

    first = json.loads(gzip.open(tmp_path / "predictions.jsonl.gz", "rt").readline())
    assert list(first) == ["prediction_arguments", "corpus_type", "languages", "feature_set", "indexer_state", "classifiers", "label_inventories", "package_version", "format_version"]
    assert first["feature_set"] == "phoible" and first["format_version"] == [1, 1, 0]
    edits = predictions.UtteranceEdits("de", "utt0", {"phoneme": ["k", "a"]}, {"phoneme": predictions.levensthein_substitutions(["k", "a"], ["g", "a", "s"])})
    assert edits.edit_operations["phoneme"] == [(Action.SUBSTITUTION, "k", "g"), (Action.INSERTION, "", "s")]
    assert json.loads(edits.to_json())["edit_operations"]["phoneme"] == [[2, "k", "g"], [0, "", "s"]]
    assert predictions.UtteranceEdits.from_json(edits.to_json()) == edits
    with pytest.raises(ValueError, match="Unsupported prediction format version"):
        predictions.PredictionMetaData.loads(json.dumps({**first, "format_version": [9, 0, 0]}))


def _enumerate_alignments(log_emissions: torch.Tensor, blank: int = 0):
    """Exhaustive CTC-style search with the reference's scoring (path score = sum of PROBABILITIES, merged by log-add):
    {collapsed token tuple: log sum_paths exp(sum_t p_t(path_t))} and the best path's first-frame timesteps."""
    import itertools
    import math

    probabilities = log_emissions.exp().double()
    frames, classes = probabilities.shape
    totals, best = {}, {}
    for path in itertools.product(range(classes), repeat=frames):
        score = sum(float(probabilities[t, c]) for t, c in enumerate(path))
        tokens, steps = [], []
        for t, c in enumerate(path):
            if c != blank and (t == 0 or c != path[t - 1]):
                tokens.append(c)
                steps.append(t + 1)
        key = tuple(tokens)
        totals[key] = math.log(math.exp(totals[key]) + math.exp(score)) if key in totals else score
        if key not in best or score > best[key][0]:
            best[key] = (score, steps)
    return totals, best


def test_beam_ctc_decoder_equals_exhaustive_search():
    """BeamCTCDecoder (predictions.py:210-226 -> flashlight lexicon-free decoder, restated in host C++): with a beam wide
    enough to keep every hypothesis the result must equal exhaustive enumeration of all alignments."""
    from allophant_b200 import predictions

    generator = torch.Generator().manual_seed(4)
    for frames, classes in [(1, 2), (3, 3), (5, 3), (4, 4), (6, 2)]:
        log_emissions = torch.log_softmax(2.0 * torch.randn(2, frames, classes, generator=generator), -1)
        lengths = torch.tensor([frames, max(1, frames - 1)])
        decoder = predictions._ctc_decoder([f"c{i}" for i in range(1, classes)], beam_width=500, n_best=4)
        assert isinstance(decoder, predictions.BeamCTCDecoder)
        results = decoder(log_emissions, lengths)
        assert len(results) == 2
        for sequence, hypotheses in enumerate(results):
            totals, best = _enumerate_alignments(log_emissions[sequence, : int(lengths[sequence])])
            ranked = sorted(totals.items(), key=lambda item: -item[1])
            assert len(hypotheses) == min(4, len(ranked))
            for hypothesis, (tokens, score) in zip(hypotheses, ranked):
                assert tuple(hypothesis.tokens.tolist()) == tokens
                assert abs(hypothesis.score - score) < 1e-9 * max(1.0, abs(score))
                assert hypothesis.words == [] and hypothesis.timesteps.dtype == torch.int32
                assert len(hypothesis.timesteps) == len(tokens)
            # 1-based first frames of the labels along the kept back pointers (the better-scoring parent at every merge)
            steps = hypotheses[0].timesteps.tolist()
            assert steps == sorted(set(steps)) and all(1 <= step <= int(lengths[sequence]) for step in steps)
    # a narrow beam still returns valid, score-sorted hypotheses; beam 1 is the greedy decoder like in the reference
    log_emissions = torch.log_softmax(torch.randn(3, 40, 6, generator=generator), -1)
    narrow = predictions.BeamCTCDecoder(["<blank>"] + list("abcde"), 5, 3)(log_emissions, torch.tensor([40, 17, 1]))
    assert all(len(h) == 3 and h[0].score >= h[1].score >= h[2].score for h in narrow[:2])
    assert all(int(t.max()) <= length for h, length in zip(narrow, (40, 17, 1)) for t in [h[0].timesteps] if len(t))
    assert isinstance(predictions._ctc_decoder(["a"], 1, 1), predictions.GreedyCTCDecoder)
    with pytest.raises(AssertionError, match="N-best can not exceed beam width"):
        predictions._ctc_decoder(["a"], 2, 3)


def test_max_pool_layers_are_rejected_like_the_reference_fails():
    """``max_pool`` layers of the sequential frontend (``frontend.py:258-259``): the unmodified reference declares more frames than
    the pool produces and its transformer then fails on the key-padding mask (golden record made by
    ``oracle/make_golden_transformer.py::max_pool_reference_behaviour``).  Here the configuration is rejected when the model is built."""
    import json
    import os

    from allophant_b200.config import MaxPoolingConfig, SequentialFrontendConfig
    from allophant_b200.network.frontend import SequentialFrontend

    with open(os.path.join(os.path.dirname(__file__), "golden", "max_pool_reference_behaviour.json")) as file:
        record = json.load(file)
    assert record["frontend_output_frames"] == 15 and min(record["frontend_declared_lengths"]) >= 10 and max(record["frontend_declared_lengths"]) == 31
    assert record["transformer_model"]["raised"] == "AssertionError" and "key_padded_mask" in record["transformer_model"]["message"]
    with pytest.raises(NotImplementedError, match="max_pool"):
        SequentialFrontend.from_config(SequentialFrontendConfig([MaxPoolingConfig(2)]), 8)
