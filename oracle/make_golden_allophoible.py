"""TEST INFRASTRUCTURE — golden vectors for restoring a reference-written ``phonetic_indexer_state``.

Runs the UNMODIFIED reference (``/root/reference/allophant/phonetic_features.py``) under the import shim on a SYNTHETIC
table in the Allophoible / PHOIBLE column layout (the real ``allophoible.csv`` is not part of the reference checkout):

1. the training-time construction ``PhoneticAttributeIndexer.from_config(config, table, LanguageInventories)``
   (``phonetic_features.py:739-786``, what ``run.py`` does before training),
2. ``indexer.state()`` — the ``PhoneticIndexerState`` a checkpoint stores (``estimator.py:214``, ``phonetic_features.py:727-728``),
   including ``table_file = original_feature_table.to_csv()`` (``:647``),
3. the restore-time construction ``PhoneticAttributeIndexer.from_config(config, state_dict=state)`` that
   ``Estimator.restore`` performs (``estimator.py:1110-1112``),

and freezes the state plus everything the model and the README usage read from the restored indexer in
``tests/golden/allophoible_restore.json``.  ``tests/test_allophoible_restore.py`` feeds the same state to
``allophant_b200.phonetic_features.PhoneticAttributeIndexer.from_config`` (a pandas-free restatement) and compares.

``langcodes`` is not installed: ``LanguageCode.from_str`` is replaced by a stand-in that knows the codes of the synthetic
table (the product carries its own ISO 639 table).  Run here only: ``python oracle/make_golden_allophoible.py``.
"""
from __future__ import annotations

import io
import json
import os
import sys
import types
import warnings
from typing import Any, Dict, List

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import reference_shim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "allophoible_restore.json")

PHOIBLE_FEATURES = [
    "tone", "stress", "syllabic", "short", "long", "consonantal", "sonorant", "continuant", "delayedRelease", "approximant", "tap", "trill",
    "nasal", "lateral", "labial", "round", "labiodental", "coronal", "anterior", "distributed", "strident", "dorsal", "high", "low", "front",
    "back", "tense", "retractedTongueRoot", "advancedTongueRoot", "periodicGlottalSource", "epilaryngealSource", "spreadGlottis",
    "constrictedGlottis", "fortis", "lenis", "raisedLarynxEjective", "loweredLarynxImplosive", "click",
]  # fmt: skip
META = ["InventoryID", "Glottocode", "ISO6393", "LanguageName", "SpecificDialect", "GlyphID", "Phoneme", "Allophones", "Marginal", "SegmentClass", "Source"]

# phones of the synthetic database: simple segments, complex ones (contours: several values per feature), a tone
SEGMENTS = ["a", "e", "i", "o", "u", "ə", "ɛ", "ɔ", "aː", "iː", "ai̯", "au̯", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "ɲ", "f", "v", "s", "z", "ʃ",
            "ʒ", "x", "h", "l", "ʎ", "r", "ɾ", "j", "w", "t̠ʃ", "d̠ʒ", "ts", "pʰ", "tʰ", "kʰ", "β", "ð", "ɣ", "ʔ", "˥", "˩"]  # fmt: skip
CONTOUR_SEGMENTS = {"ai̯", "au̯", "t̠ʃ", "d̠ʒ", "ts"}


def synthetic_features(rng: np.random.Generator) -> Dict[str, Dict[str, str]]:
    """phone -> feature -> value ("+", "-", "0" or a comma separated contour such as "-,+")"""
    table = {}
    for phone in SEGMENTS:
        row = {}
        for feature in PHOIBLE_FEATURES:
            if feature == "tone":
                row[feature] = "+" if phone in ("˥", "˩") else "0"
            elif phone in CONTOUR_SEGMENTS and rng.random() < 0.25:
                row[feature] = ",".join(rng.choice(["+", "-"], size=2))
            else:
                row[feature] = str(rng.choice(["+", "-", "0"], p=[0.35, 0.5, 0.15]))
        table[phone] = row
    return table


def synthetic_database(seed: int = 7) -> str:
    """CSV text with PHOIBLE's columns: several inventories per language (sources / dialects), marginal phonemes, rows
    without allophones, allophones that are no phoneme of any selected inventory, and a preferred dialect (``eng``)."""
    import csv

    rng = np.random.default_rng(seed)
    features = synthetic_features(rng)
    inventories = [
        # (InventoryID, Glottocode, ISO6393, LanguageName, SpecificDialect, Source, phonemes)
        (1, "span1234", "spa", "Spanish", None, "spa", ["a", "e", "i", "o", "u", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ɲ", "f", "s", "x", "l", "ʎ", "r", "ɾ", "j", "w", "t̠ʃ", "β", "ð", "ɣ"]),
        (2, "span1234", "spa", "Spanish", "Castilian", "upsid", ["a", "e", "i", "o", "u", "p", "b", "t", "d", "k", "ɡ", "m", "n", "s", "l", "r"]),
        (3, "ital1282", "ita", "Italian", None, "spa", ["a", "e", "i", "o", "u", "ɛ", "ɔ", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ɲ", "f", "v", "s", "z", "ʃ", "l", "ʎ", "r", "j", "w", "ts", "t̠ʃ", "d̠ʒ"]),
        (4, "stan1293", "eng", "English", "Western and Mid-Western US; Southern California", "uz", ["a", "e", "i", "o", "u", "ə", "aː", "iː", "ai̯", "au̯", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "f", "v", "s", "z", "ʃ", "ʒ", "h", "l", "r", "j", "w", "t̠ʃ", "d̠ʒ"]),
        (5, "stan1293", "eng", "English", "Received Pronunciation", "spa", ["a", "e", "i", "o", "u", "ə", "ɛ", "ɔ", "aː", "iː", "ai̯", "au̯", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "f", "v", "s", "z", "ʃ", "ʒ", "h", "l", "r", "j", "w", "t̠ʃ", "d̠ʒ", "ʔ"]),
        (6, "mand1415", "cmn", "Mandarin Chinese", None, "spa", ["a", "i", "u", "ə", "p", "pʰ", "t", "tʰ", "k", "kʰ", "m", "n", "ŋ", "f", "s", "x", "l", "ts", "˥", "˩"]),
        (7, "germ1287", "deu", "German", None, "upsid", ["a", "e", "i", "o", "u", "ə", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "f", "v", "s", "z", "ʃ", "x", "h", "l", "r", "j", "ts"]),
    ]  # fmt: skip
    allophone_variants = {"b": ["b", "β"], "d": ["d", "ð"], "ɡ": ["ɡ", "ɣ"], "t": ["t", "tʰ"], "k": ["k", "kʰ"], "n": ["n", "ŋ"], "r": ["r", "ɾ"], "e": ["e", "ɛ"], "o": ["o", "ɔ"]}
    buffer = io.StringIO()
    writer = csv.writer(buffer, lineterminator="\n")
    writer.writerow(META + PHOIBLE_FEATURES)
    glyph = 0
    for inventory_id, glotto, iso, name, dialect, source, phonemes in inventories:
        for position, phoneme in enumerate(phonemes):
            glyph += 1
            allophones = [phoneme]
            if phoneme in allophone_variants and rng.random() < 0.7:
                allophones = allophone_variants[phoneme]
            marginal = ""
            if source == "upsid":
                allophone_cell = ""  # UPSID rows carry no allophone information
            else:
                allophone_cell = " ".join(allophones)
                if position % 11 == 10:
                    marginal = "TRUE"
                elif position % 3 == 0:
                    marginal = "FALSE"
            segment_class = "tone" if phoneme in ("˥", "˩") else ("vowel" if features[phoneme]["syllabic"].startswith("+") else "consonant")
            writer.writerow(
                [inventory_id, glotto, iso, name, dialect or "", f"{glyph:04X}", phoneme, allophone_cell, marginal, segment_class, source]
                + [features[phoneme][feature] for feature in PHOIBLE_FEATURES]
            )
    return buffer.getvalue()


def install_language_codes() -> None:
    """Stand-in for ``allophant.language_codes.LanguageCode.from_str`` (langcodes is absent): ISO 639-1/-3 of the synthetic table."""
    import importlib

    module = importlib.import_module("allophant.language_codes")
    alpha3 = {"es": "spa", "it": "ita", "en": "eng", "de": "deu", "zh": "zho"}

    def from_str(cls, language_code: str, standardize: bool = False, macro: bool = False):
        code = language_code.lower().split("-")[0]
        code = alpha3.get(code, code)
        if macro and code == "cmn":
            code = "zho"
        return cls(code, code, code, None)

    module.LanguageCode.from_str = classmethod(from_str)


def install_pandas2_groupby_apply() -> None:
    """The reference pins pandas 2; pandas 3 (installed here) no longer passes the grouping column to the function given to
    ``DataFrameGroupBy.apply`` (``include_groups=True`` was removed), which ``_filter_inventory`` (phonetic_features.py:1045-1064)
    relies on.  This restores the pandas 2 behaviour for that call: groups in sorted key order, grouping column included,
    ``group.name`` set, results concatenated under the group keys.  A third-party compatibility patch — the reference is untouched."""
    import pandas as pd
    from pandas.core.groupby import DataFrameGroupBy

    def apply(self, func, *args, **kwargs):
        kwargs.pop("include_groups", None)
        pieces, keys = [], []
        for name, group in self:
            group = group.copy()
            object.__setattr__(group, "name", name)
            pieces.append(func(group, *args, **kwargs))
            keys.append(name)
        key_names = self.keys if isinstance(self.keys, list) else [self.keys]
        return pd.concat(pieces, keys=keys, names=key_names)

    DataFrameGroupBy.apply = apply
    # pandas 2 semantics for ``dtype=str`` (object columns; the reference assigns arrays into them, phonetic_features.py:554)
    pd.options.future.infer_string = False


def tolist(value: Any) -> Any:
    if hasattr(value, "tolist"):
        return value.tolist()
    return list(value)


def describe(indexer: Any, inventories: Dict[str, List[str]]) -> Dict[str, Any]:
    """Everything ``Allophant.from_config`` (acoustic_model.py:405-446, 986-1003), ``Estimator.predict`` callers and the README read."""
    shared = indexer.allophone_data.shared_phone_indexer
    languages = indexer.language_allophones
    out: Dict[str, Any] = {
        "phonemes": tolist(indexer.phonemes),
        "feature_names": list(indexer.feature_names),
        "feature_categories": {name: list(indexer.feature_categories(name)) for name in indexer.feature_names},
        "sizes": {name: indexer.size(name) for name in indexer.feature_names},
        "total_size": indexer.size(),
        "composition_features": list(indexer.composition_features),
        "full_phonemes": tolist(indexer.full_attributes.phonemes),
        "full_feature_names": list(indexer.full_attributes.feature_names),
        "full_dense": indexer.full_attributes.dense_feature_table.long().tolist(),
        "full_categories": {name: list(indexer.full_attributes.feature_categories(name)) for name in indexer.full_attributes.feature_names},
        "subset_dense": indexer.attributes.dense_feature_table.long().tolist(),
        "shared_phonemes": tolist(shared.phonemes),
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
        "contours": {phoneme: [tolist(v) for v in indexer.full_attributes.feature_vector(phoneme)] for phoneme in ("ai̯", "t̠ʃ", "a")},
    }
    # README usage: ``attribute_indexer.attributes.subset(inventory)`` (phonemes of the training subset) and the same over
    # every phone of the database (``full_subset_attributes``, evaluation.py)
    out["custom_subset_full"] = {
        "phonemes": tolist(indexer.full_subset_attributes.subset(inventories["custom"]).phonemes),
        "feature_names": list(indexer.full_subset_attributes.subset(inventories["custom"]).feature_names),
        "dense": indexer.full_subset_attributes.subset(inventories["custom"]).dense_feature_table.long().tolist(),
    }
    subset = indexer.attributes.subset(inventories["trained_subset"])
    out["custom_subset"] = {
        "phonemes": tolist(subset.phonemes),
        "feature_names": list(subset.feature_names),
        "dense": subset.dense_feature_table.long().tolist(),
        "phoneme_categories": list(subset.feature_categories("phoneme")),
    }
    return out


def main() -> None:
    ref = reference_shim.reference_modules()
    install_language_codes()
    install_pandas2_groupby_apply()
    pf, cfg = ref.phonetic_features, ref.config
    table = synthetic_database()
    attributes = ["stress", "syllabic", "consonantal", "sonorant", "continuant", "nasal", "labial", "coronal", "dorsal", "high", "low", "front", "back"]
    entries = [cfg.ProjectionEntryConfig("phoneme", ["OUTPUT", *attributes])] + [cfg.ProjectionEntryConfig(name, ["OUTPUT"]) for name in attributes]
    projection = cfg.ProjectionConfig(entries, phoneme_layer=cfg.PhonemeLayerType.ALLOPHONES, feature_set=cfg.FeatureSet.PHOIBLE)
    config = types.SimpleNamespace(nn=types.SimpleNamespace(projection=projection))

    # training-time inventories (what the corpus G2P produced, run.py): per-language phoneme lists; "it" includes a phoneme its
    # database inventory lacks ("x": filled from another inventory by _filter_inventory, phonetic_features.py:1045-1064)
    training = pf.LanguageInventories(
        {
            0: ["a", "e", "i", "o", "u", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ɲ", "f", "s", "x", "l", "r", "j", "t̠ʃ"],
            1: ["a", "e", "i", "o", "u", "ɛ", "ɔ", "p", "b", "t", "d", "k", "ɡ", "m", "n", "f", "v", "s", "z", "ʃ", "l", "r", "ts", "x"],
            2: ["a", "i", "u", "ə", "aː", "ai̯", "p", "b", "t", "d", "k", "ɡ", "m", "n", "ŋ", "f", "v", "s", "z", "h", "l", "r", "w", "d̠ʒ"],
        },
        ["es", "it", "en"],
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        trained = pf.PhoneticAttributeIndexer.from_config(config, io.StringIO(table), training)
        state = trained.state()
        restored = pf.PhoneticAttributeIndexer.from_config(config, state_dict=state)

    inventories = {
        "es": restored.phoneme_inventory("es"),
        "custom": ["a", "ai̯", "au̯", "b", "e", "f", "ɡ", "l", "ʎ", "m", "ɲ", "o", "p", "ɾ", "s", "t̠ʃ"],
        "unseen": ["pʰ", "tʰ", "kʰ", "˥", "ʔ", "ɣ"],
        "trained_subset": ["a", "ai̯", "b", "e", "f", "ɡ", "l", "m", "ɲ", "o", "p", "s", "t̠ʃ"],
    }
    golden = {
        "state": {
            "phoneme_inventory": list(state.phoneme_inventory),
            "language_allophones": {
                "allophones": {str(l): {str(p): list(map(int, q)) for p, q in m.items()} for l, m in state.language_allophones.allophones.items()},
                "languages": list(state.language_allophones.languages),
                "shared_phones": list(state.language_allophones.shared_phones),
            },
            "table_file": state.table_file,
        },
        "classes": [[entry.name, list(entry.dependencies)] for entry in entries],
        "inventories": inventories,
        "trained": describe(trained, inventories),
        "restored": describe(restored, inventories),
    }
    with open(GOLDEN, "w", encoding="utf-8") as file:
        json.dump(golden, file, ensure_ascii=False, indent=0, sort_keys=True)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes;", len(golden["restored"]["phonemes"]), "phonemes,", len(golden["restored"]["shared_phonemes"]), "shared phones")


if __name__ == "__main__":
    main()
