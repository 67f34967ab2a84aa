"""Shared test helpers: oracle construction (CPU) and the CUDA model loaded with the oracle's weights."""
from __future__ import annotations

import os
from functools import lru_cache
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from oracle import restatement

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str) -> Dict[str, Any]:
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def synthetic_allophones(n_languages: int, n_phonemes: int, n_phones: int, seed: int):
    rng = np.random.default_rng(seed)
    result = {}
    for language in range(n_languages):
        inventory = sorted(rng.choice(n_phonemes, size=max(4, n_phonemes // 2), replace=False).tolist())
        mapping = {}
        for phoneme in inventory:
            count = int(rng.integers(1, 4))
            mapping[int(phoneme)] = sorted(int(p) for p in rng.choice(n_phones, size=count, replace=False))
        result[language] = mapping
    return result


def spec_for_case(case: Dict[str, Any]) -> restatement.OracleSpec:
    """Same construction as oracle/make_golden.py:build_spec (kept in sync by test_oracle_golden)."""
    spec = restatement.multitask_spec(**case["spec"])
    allophones = case.get("allophones")
    if allophones is not None:
        n_phonemes = case["spec"]["n_train_phonemes"]
        spec.allophones = synthetic_allophones(allophones["n_languages"], n_phonemes, allophones["n_phones"], allophones["seed"])
        spec.n_phones = allophones["n_phones"]
        rng = np.random.default_rng(77)
        table = rng.integers(0, 3, size=(spec.n_phones, len(restatement.PHOIBLE_FEATURES)))
        table[:3, :] = np.arange(3)[:, None]
        spec.feature_table = table
    return spec


def batch_for_case(fixture: Dict[str, Any]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lengths = fixture["lengths"]
    audio = restatement.synthetic_audio(len(lengths), int(lengths.max()), seed=0)
    audio = audio * restatement.mask_sequence(lengths)
    return audio, lengths, fixture["language_ids"]


def cuda_model_for_spec(
    spec: restatement.OracleSpec, oracle: Any, device: str = "cuda", acoustic_config: Any = None, feature_size: int = 1, state_dict: Any = None,
    time_layers: Any = None, extra_dependencies: Any = None,
):
    """Builds allophant_b200's Allophant with the architecture of ``spec`` and loads the oracle's weights (or ``state_dict``);
    ``acoustic_config`` replaces the wav2vec2 encoder (the from-scratch transformer cases)."""
    from allophant_b200.attribute_graph import AttributeGraph, AttributeNode
    from allophant_b200.config import (
        Architecture,
        CTCLossConfig,
        EmbeddingCompositionConfig,
        PhonemeLayerType,
        ProjectionConfig,
        ProjectionEntryConfig,
        Wav2Vec2PretrainedConfig,
    )
    from allophant_b200.network import wav2vec2
    from allophant_b200.network.acoustic_model import Allophant
    from allophant_b200.phonetic_features import AllophoneData, ArticulatoryAttributes, LanguageAllophoneMappings, PhoneticAttributeIndexer

    features = restatement.PHOIBLE_FEATURES
    table = np.asarray(spec.feature_table)
    categories = {f: [str(v) for v in range(int(table[:, i].max()) + 1)] for i, f in enumerate(features)}
    phoneme_class = next(c for c in spec.classes if c.name == "phoneme")
    if spec.allophones is None:
        names = [f"p{i}" for i in range(table.shape[0])]
        attributes = ArticulatoryAttributes(names, features, table, categories)
        indexer = PhoneticAttributeIndexer(attributes, names, features, features + ["phoneme"])
        phoneme_layer = PhonemeLayerType.SHARED
    else:
        phones = [f"ph{i}" for i in range(spec.n_phones)]
        shared = ArticulatoryAttributes(phones, features, table, categories)
        mappings = LanguageAllophoneMappings(spec.allophones, [f"l{i}" for i in range(len(spec.allophones))], phones)
        indexer = PhoneticAttributeIndexer(
            shared, [f"p{i}" for i in range(phoneme_class.size)], features, features + ["phoneme"], mappings, AllophoneData(shared)
        )
        phoneme_layer = PhonemeLayerType.ALLOPHONES
    from allophant_b200.config import MultiheadAttentionConfig

    def time_layer(name):
        options = (time_layers or {}).get(name)
        return None if options is None else MultiheadAttentionConfig(options["num_heads"], options["positional_embeddings"])

    def dependencies_of(c):
        return list(c.dependencies) + list((extra_dependencies or {}).get(c.name, []))

    projection = ProjectionConfig(
        [ProjectionEntryConfig(c.name, dependencies_of(c), time_layer(c.name)) for c in spec.classes],
        phoneme_layer=phoneme_layer,
        acoustic_model_dropout=0.2,
        dependency_blanks=spec.dependency_blanks,
        embedding_composition=None if spec.embedding_size is None else EmbeddingCompositionConfig(spec.embedding_size),
    )
    model_id = "facebook/wav2vec2-xls-r-300m"
    if spec.encoder_overrides:
        import dataclasses

        model_id = "test/" + "-".join(f"{k}{v}" for k, v in sorted(spec.encoder_overrides.items()))
        wav2vec2.KNOWN_MODELS[model_id] = dataclasses.replace(wav2vec2.KNOWN_MODELS["facebook/wav2vec2-xls-r-300m"], **spec.encoder_overrides)
    architecture = Architecture(
        16_000_000, projection, Wav2Vec2PretrainedConfig(model_id) if acoustic_config is None else acoustic_config, loss=CTCLossConfig()
    )
    graph = AttributeGraph(AttributeNode(c.name, c.size, time_layer(c.name), dependencies_of(c)) for c in spec.classes)
    model = Allophant.from_config(architecture, feature_size, 16000, graph, indexer, load_pretrained_weights=False)
    result = model.load_state_dict(oracle.state_dict() if state_dict is None else state_dict, strict=True)
    assert not result.missing_keys and not result.unexpected_keys
    return model.to(device).eval(), indexer


def rel_err(value: torch.Tensor, reference: torch.Tensor) -> float:
    return float((value.double().cpu() - reference.double().cpu()).abs().max() / reference.double().abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------ train()-mode masks (restated from aph_common.cuh)
def _fmix(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def keep_mask(dropout, rows: int, cols: int) -> torch.Tensor:
    """fp32 [rows, cols] multiplicative mask (scale or 0) of ``allophant_b200.ops.Dropout`` — the numpy restatement of
    ``drop_row_key`` / ``drop_hash`` / ``drop_keep`` (``aph_common.cuh``)."""
    if dropout.threshold == 0:
        return torch.ones(rows, cols)
    row = np.arange(rows, dtype=np.uint64)[:, None]
    col = np.arange(cols, dtype=np.uint64)[None, :]
    key = _fmix(np.uint64(dropout.seed) ^ ((row * 0x9E3779B1) & 0xFFFFFFFF))
    hashed = _fmix(key ^ (((col >> 1) * 0x9E3779B1) & 0xFFFFFFFF))
    half = np.where(col & 1, hashed >> 16, hashed & 0xFFFF)
    return torch.from_numpy((half >= dropout.threshold).astype(np.float32) * np.float32(dropout.scale))


def regularisation_masks(stochastic, n_utt: int, seq: int, hidden: int, heads: int, n_layers: int, skipped, spec_mask=None, intermediate: int = 0):
    """The explicit masks ``oracle.restatement.OracleModel.explicit_regularisation`` takes, equal to what the CUDA
    kernels derive from ``stochastic`` (an ``allophant_b200.engine.Stochastic``)."""
    rows = n_utt * seq
    masks = {
        "feature_projection": keep_mask(stochastic.feature_projection(), rows, hidden).view(n_utt, seq, hidden),
        "encoder_input": keep_mask(stochastic.encoder_input(), rows, hidden).view(n_utt, seq, hidden),
        "skip": list(skipped),
    }
    for layer in range(n_layers):
        masks[f"attention.{layer}"] = keep_mask(stochastic.attention(layer), n_utt * heads * seq, seq).view(n_utt, heads, seq, seq)
        masks[f"attention_output.{layer}"] = keep_mask(stochastic.attention_output(layer), rows, hidden).view(n_utt, seq, hidden)
        masks[f"feed_forward_output.{layer}"] = keep_mask(stochastic.feed_forward_output(layer), rows, hidden).view(n_utt, seq, hidden)
        if intermediate and stochastic.activation(layer).threshold:
            masks[f"activation.{layer}"] = keep_mask(stochastic.activation(layer), rows, intermediate).view(n_utt, seq, intermediate)
    if spec_mask is not None:
        masks["spec"] = spec_mask.view(n_utt, seq).bool()
    return masks


def transformer_model_for_golden(golden: Dict[str, Any], device: str = "cuda"):
    """allophant_b200's Allophant over the from-scratch transformer encoder of a ``tests/golden/transformer_*.pt`` case
    (same construction as oracle/make_golden_transformer.py), loaded with the REFERENCE's state_dict."""
    from allophant_b200.config import TransformerAcousticModelConfig

    case = golden["case"]
    spec = restatement.multitask_spec(**case["spec"])
    spec.embedding_size = 64
    if case["output_layers"]:
        first = spec.classes[0]
        spec.classes[0] = restatement.ClassSpec(first.name, first.size, list(first.dependencies) + [f"OUTPUT_{i}" for i in case["output_layers"]])
    options = case["acoustic"]
    mapping = dict(
        type="pre-ln-transformer",
        transformer=options["transformer"],
        frontend=options["frontend"],
        sequential_frontend=None if options["sequential_frontend"] is None else {"layers": options["sequential_frontend"]},
        elementwise_affine=options["elementwise_affine"],
    )
    config = TransformerAcousticModelConfig.load(mapping)
    return cuda_model_for_spec(spec, None, device, config, case["feature_size"], golden["state_dict"])
