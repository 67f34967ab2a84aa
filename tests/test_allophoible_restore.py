"""Restoring a reference-written ``phonetic_indexer_state`` (embedded Allophoible CSV) without pandas.

The golden file comes from the UNMODIFIED reference (``oracle/make_golden_allophoible.py``): its training-time indexer, the
state it stores in a checkpoint, and what its restore-time indexer answers.  ``allophant_b200`` must rebuild the same tables
from that state (``allophant/phonetic_features.py:601-786``, ``estimator.py:1086-1126``)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from allophant_b200 import allophoible
from allophant_b200.config import FeatureSet, PhonemeLayerType, ProjectionConfig, ProjectionEntryConfig
from allophant_b200.language_codes import LanguageCode, standardize_to_iso6393
from allophant_b200.phonetic_features import LanguageAllophoneMappings, LanguageInventories, PhoneticAttributeIndexer

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "allophoible_restore.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN, encoding="utf-8") as file:
        return json.load(file)


def make_config(golden):
    entries = [ProjectionEntryConfig(name, list(dependencies)) for name, dependencies in golden["classes"]]
    projection = ProjectionConfig(entries, phoneme_layer=PhonemeLayerType.ALLOPHONES, feature_set=FeatureSet.PHOIBLE)
    return types.SimpleNamespace(nn=types.SimpleNamespace(projection=projection))


def describe(indexer, inventories):
    shared = indexer.allophone_data.shared_phone_indexer
    languages = indexer.language_allophones
    custom = indexer.attributes.subset(inventories["trained_subset"])
    custom_full = indexer.full_subset_attributes.subset(inventories["custom"])
    return {
        "phonemes": indexer.phonemes.tolist(),
        "feature_names": list(indexer.feature_names),
        "feature_categories": {name: list(indexer.feature_categories(name)) for name in indexer.feature_names},
        "sizes": {name: indexer.size(name) for name in indexer.feature_names},
        "total_size": indexer.size(),
        "composition_features": list(indexer.composition_features),
        "full_phonemes": indexer.full_attributes.phonemes.tolist(),
        "full_feature_names": list(indexer.full_attributes.feature_names),
        "full_dense": indexer.full_attributes.dense_feature_table.long().tolist(),
        "full_categories": {name: list(indexer.full_attributes.feature_categories(name)) for name in indexer.full_attributes.feature_names},
        "subset_dense": indexer.attributes.dense_feature_table.long().tolist(),
        "shared_phonemes": shared.phonemes.tolist(),
        "shared_feature_names": list(shared.feature_names),
        "shared_dense": shared.dense_feature_table.long().tolist(),
        "language_allophones": {
            "allophones": {str(l): {str(p): list(map(int, q)) for p, q in m.items()} for l, m in languages.allophones.items()},
            "languages": list(languages.languages),
            "shared_phones": list(languages.shared_phones),
        },
        "phone_categories": list(indexer.feature_categories("phone")),
        "phoneme_inventory": {code: indexer.phoneme_inventory(code) for code in ("es", "it", "eng", "cmn", "de")},
        "phoneme_inventory_union": indexer.phoneme_inventory(["es", "it"]),
        "composition_matrix": {name: indexer.composition_feature_matrix(inventory).tolist() for name, inventory in inventories.items()},
        "contours": {phoneme: [v.tolist() for v in indexer.full_attributes.feature_vector(phoneme)] for phoneme in ("ai̯", "t̠ʃ", "a")},
        "custom_subset": {
            "phonemes": custom.phonemes.tolist(),
            "feature_names": list(custom.feature_names),
            "dense": custom.dense_feature_table.long().tolist(),
            "phoneme_categories": list(custom.feature_categories("phoneme")),
        },
        "custom_subset_full": {
            "phonemes": custom_full.phonemes.tolist(),
            "feature_names": list(custom_full.feature_names),
            "dense": custom_full.dense_feature_table.long().tolist(),
        },
    }


def assert_same(ours, expected):
    assert set(ours) == set(expected)
    for key in expected:
        assert ours[key] == expected[key], key


def test_restore_from_a_reference_written_state(golden):
    indexer = PhoneticAttributeIndexer.from_config(make_config(golden), state_dict=golden["state"])
    assert_same(describe(indexer, golden["inventories"]), golden["restored"])
    # the state written back is the one that was read (checkpoints stay interchangeable)
    state = indexer.state()
    assert state["table_file"] == golden["state"]["table_file"]
    assert state["phoneme_inventory"] == golden["state"]["phoneme_inventory"]
    assert {str(l): {str(p): q for p, q in m.items()} for l, m in state["language_allophones"]["allophones"].items()} == golden["state"][
        "language_allophones"
    ]["allophones"]


def test_training_time_construction_matches_the_reference(golden):
    """The path ``run.py`` takes before training: corpus inventories in, allophone mappings derived from the database."""
    reference_state = golden["state"]
    spa = ["a", "e", "i", "o", "u", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ɲ", "f", "s", "x", "l", "r", "j", "t̠ʃ"]
    ita = ["a", "e", "i", "o", "u", "ɛ", "ɔ", "p", "b", "t", "d", "k", "ɡ", "m", "n", "f", "v", "s", "z", "ʃ", "l", "r", "ts", "x"]
    eng = ["a", "i", "u", "ə", "aː", "ai̯", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "f", "v", "s", "z", "h", "l", "r", "w", "d̠ʒ"]
    training = LanguageInventories({0: spa, 1: ita, 2: eng}, ["es", "it", "en"])
    # the CSV of the state is the original table with "Phoneme" moved to the front: either layout is accepted
    indexer = PhoneticAttributeIndexer.from_config(make_config(golden), reference_state["table_file"], training)
    assert_same(describe(indexer, golden["inventories"]), golden["trained"])


def test_estimator_restore_accepts_the_reference_layout(golden):
    """``Estimator.restore`` on a checkpoint dictionary in the reference's layout whose indexer state embeds the database CSV
    (``estimator.py:1086-1126``): the model is built from the restored indexer and its parameters load."""
    from allophant_b200.attribute_graph import AttributeGraph, AttributeNode
    from allophant_b200.config import Config
    from allophant_b200.network.acoustic_model import Allophant

    import dataclasses

    from allophant_b200.network import wav2vec2

    model_id = "test/restore-1-layer"
    wav2vec2.KNOWN_MODELS[model_id] = dataclasses.replace(wav2vec2.KNOWN_MODELS["facebook/wav2vec2-xls-r-300m"], num_hidden_layers=1)
    config = Config.load(
        {
            "nn": {
                "batch_size": 8,
                "acoustic_model": {"type": "wav2vec2-pretrained", "model_id": model_id},
                "projection": {
                    "classes": [{"name": name, "dependencies": dependencies} for name, dependencies in golden["classes"]],
                    "phoneme_layer": "allophones",
                    "embedding_composition": {"embedding_size": 32},
                },
            }
        }
    )
    indexer = PhoneticAttributeIndexer.from_config(config, state_dict=golden["state"])
    n_phones = len(golden["state"]["language_allophones"]["shared_phones"])
    graph = AttributeGraph(
        AttributeNode(entry.name, n_phones if entry.name == "phoneme" else indexer.size(entry.name), entry.time_layer, list(entry.dependencies))
        for entry in config.nn.projection.classes
    )
    torch.manual_seed(0)
    model = Allophant.from_config(config.nn, 1, 16000, graph, indexer, load_pretrained_weights=False)
    checkpoint = {
        "config": config.dump(),
        "allophant_version": "1.0.0",
        "feature_size": 1,
        "sample_rate": 16000,
        "attribute_graph": graph.state(),
        "epoch": {"epoch": 3, "global_step": 70, "step": 10},
        "phonetic_indexer_state": golden["state"],
        "dataset_meta_data": [],
        "model_state": {name: value.clone() for name, value in model.state_dict().items()},
        "additional": None,
        "history": [],
        "optimization_states": None,
    }
    from allophant_b200.estimator import Estimator

    estimator, restored = Estimator.restore(checkpoint, device="cpu")
    assert restored.phonemes.tolist() == golden["restored"]["phonemes"]
    assert restored.allophone_data.shared_phone_indexer.phonemes.tolist() == golden["restored"]["shared_phonemes"]
    assert estimator.epoch["global_step"] == 70
    for name, value in estimator.model.state_dict().items():
        assert torch.equal(value, checkpoint["model_state"][name]), name
    # the composed-embedding table follows the shared phones' categories: same sizes as the reference would allocate
    composition = estimator.model._projection._layers["phoneme"]._composition_layer
    dense = torch.tensor(golden["restored"]["shared_dense"])
    assert composition._attribute_embeddings.weight.shape[0] == int((dense.max(0).values + 1).sum()) + 1


def test_zero_phoneme_removal_and_macro_language_fallback():
    header = "InventoryID,Glottocode,ISO6393,LanguageName,SpecificDialect,GlyphID,Phoneme,Allophones,Marginal,SegmentClass,Source,tone,stress,syllabic"
    rows = [
        "1,mand1415,cmn,Mandarin,,0001,a,a ∅,,vowel,spa,0,-,+",
        "1,mand1415,cmn,Mandarin,,0002,p,p pʰ,,consonant,spa,0,-,-",
        "1,mand1415,cmn,Mandarin,,0003,i,i ∅,TRUE,vowel,spa,0,-,+",
        "2,stan1295,deu,German,,0004,pʰ,pʰ,,consonant,spa,0,\"-,+\",-",
    ]
    table = allophoible.read_allophoible("\n".join([header, *rows]) + "\n")
    with pytest.warns(allophoible.LanguageMappingWarning):
        selected = allophoible.extract_allophone_inventories(table, ["zh"], None, prefer_default_dialects=True, remove_zero_phoneme=True)
    inventories = allophoible.allophone_inventories(table, selected)
    # "zh" has no inventory of its own: the Mandarin one is used and relabelled; the aspirated allophone comes first (InventoryID 0)
    assert inventories.phonemes == ["pʰ", "a", "p", "i"]
    assert inventories.iso6393 == [None, "zho", "zho", "zho"]
    assert inventories.inventory_ids == [0, 1, 1, 1]
    assert inventories.allophones == [None, ["a"], ["p", "pʰ"], ["i"]]
    with pytest.raises(ValueError, match="don't contain allophone data"):
        allophoible.extract_allophone_inventories(table, ["fr"], None)


def test_language_codes():
    assert standardize_to_iso6393("es") == "spa" and standardize_to_iso6393("eng") == "eng" and standardize_to_iso6393("en-US") == "eng"
    assert LanguageCode.from_str("ger").alpha3 == "deu" and LanguageCode.from_str("de").alpha3_b == "ger"
    assert LanguageCode.from_str("cmn", True, True).alpha3_t == "zho" and LanguageCode.from_str("cmn").alpha3 == "cmn"
    with pytest.raises(ValueError):
        LanguageCode.from_str("cmn", macro=True)
    with pytest.raises(ValueError):
        LanguageCode.from_str("x1")


def test_select_largest_inventories_prefers_the_configured_dialect_then_size():
    header = "InventoryID,Glottocode,ISO6393,LanguageName,SpecificDialect,GlyphID,Phoneme,Allophones,Marginal,SegmentClass,Source,tone"
    rows = (
        [f"1,g,eng,English,Received Pronunciation,{i},p{i},p{i},,consonant,spa,0" for i in range(5)]
        + [f"2,g,eng,English,Western and Mid-Western US; Southern California,{i},p{i},p{i},,consonant,uz,0" for i in range(3)]
        + [f"3,g,fra,French,,{i},p{i},p{i},,consonant,ph,0" for i in range(2)]
        + [f"4,g,fra,French,,{i},p{i},p{i},,consonant,spa,0" for i in range(4)]
        + [f"5,g,ita,Italian,,{i},p{i},p{i},,consonant,spa,0" for i in range(2)]
        + [f"6,g,ita,Italian,,{i},p{i},p{i},,consonant,ph,0" for i in range(2)]
    )
    table = allophoible.read_allophoible("\n".join([header, *rows]) + "\n")
    chosen = allophoible.select_largest_inventories(table.rows, table, allophoible.default_dialects())
    assert chosen == [("spa", "fra", None), ("uz", "eng", "Western and Mid-Western US; Southern California"), ("ph", "ita", None)]
    assert allophoible.select_largest_inventories(table.rows, table, None)[0] == ("spa", "eng", "Received Pronunciation")
