"""Phoneme/feature tables: the producer of ``composition_feature_matrix`` (input contract of the heads).

The reference derives these tables from the Allophoible CSV with pandas
(``allophant/phonetic_features.py:246-971``); that ingestion is outside the hot path (SURVEY.md §2
row 9) and the CSV is not redistributable here.  This module keeps the few pieces the model and
the decoders touch — ``ArticulatoryAttributes.dense_feature_table`` / ``subset`` (265-309),
``PhoneticAttributeIndexer.composition_feature_matrix`` (808-818), ``phoneme_inventory``,
``feature_categories`` and the serialisable state (40-44, 111-115) — over a plain integer table
``[phonemes, features]`` holding the first value of every feature contour.
"""
from __future__ import annotations

import io
import json
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np
import torch
from torch import Tensor


@dataclass
class LanguageAllophoneMappings:
    """language id -> {phoneme index -> [shared phone indices]} (``phonetic_features.py:40-44``)."""

    allophones: Dict[int, Dict[int, List[int]]]
    languages: List[str]
    shared_phones: List[str]


@dataclass
class PhoneticIndexerState:
    phoneme_inventory: List[str]
    language_allophones: Optional[LanguageAllophoneMappings] = None
    table_file: Optional[str] = None


class _PhonemeIndex(list):
    """List of phoneme strings with the ``.tolist()`` the reference's pandas ``Index`` offers."""

    def tolist(self) -> List[str]:
        return list(self)


class ArticulatoryAttributes:
    def __init__(self, phonemes: Sequence[str], feature_names: Sequence[str], table: np.ndarray, feature_categories: Mapping[str, Sequence[str]]):
        table = np.asarray(table, dtype=np.int64).reshape(len(phonemes), len(feature_names))
        self._phonemes = _PhonemeIndex(phonemes)
        self._positions = {phoneme: index for index, phoneme in enumerate(self._phonemes)}
        self._feature_names = list(feature_names)
        self._feature_categories = {name: list(feature_categories[name]) for name in self._feature_names}
        self._table = table
        # float tensor like the reference's (callers apply `.long()`, acoustic_model.py:194, predictions.py:241)
        self._dense_feature_table = torch.from_numpy(table.astype(np.float32))

    @property
    def dense_feature_table(self) -> Tensor:
        return self._dense_feature_table

    @property
    def phonemes(self) -> _PhonemeIndex:
        return self._phonemes

    @property
    def feature_names(self) -> List[str]:
        return self._feature_names

    def feature_categories(self, name: str) -> List[str]:
        return self._feature_categories[name]

    def phoneme_indices(self, phonemes: Iterable[str]) -> np.ndarray:
        try:
            return np.array([self._positions[phoneme] for phoneme in phonemes], dtype=np.int64)
        except KeyError as error:
            raise KeyError(f"Phoneme {error.args[0]!r} is not part of the feature table") from None

    def subset(self, phonemes: Optional[Sequence[str]] = None, attribute_subset: Optional[Sequence[str]] = None, reindex_phonemes: bool = True) -> "ArticulatoryAttributes":
        rows = np.arange(len(self._phonemes)) if phonemes is None else self.phoneme_indices(phonemes)
        names = self._feature_names if attribute_subset is None else list(attribute_subset)
        columns = [self._feature_names.index(name) for name in names]
        table = self._table[np.ix_(rows, columns)].copy()
        categories = dict(self._feature_categories)
        selected = [self._phonemes[int(r)] for r in rows]
        if reindex_phonemes and "phoneme" in names:
            table[:, names.index("phoneme")] = np.arange(len(selected))
            categories["phoneme"] = selected
        return ArticulatoryAttributes(selected, names, table, categories)

    def __len__(self) -> int:
        return len(self._phonemes)

    # -- CSV round trip (the checkpoint stores the whole table as text, phonetic_features.py:647) --
    def to_csv(self) -> str:
        buffer = io.StringIO()
        buffer.write("#allophant_b200-feature-table\t" + json.dumps(self._feature_categories, ensure_ascii=False) + "\n")
        buffer.write("phoneme\t" + "\t".join(self._feature_names) + "\n")
        for phoneme, row in zip(self._phonemes, self._table):
            buffer.write(phoneme + "\t" + "\t".join(str(int(v)) for v in row) + "\n")
        return buffer.getvalue()

    @classmethod
    def from_csv(cls, text: str) -> "ArticulatoryAttributes":
        lines = text.splitlines()
        if not lines or not lines[0].startswith("#allophant_b200-feature-table\t"):
            raise NotImplementedError(
                "this checkpoint embeds the reference's Allophoible CSV; parsing it (pandas/panphon ingestion, "
                "phonetic_features.py:601-700) is outside this build — pass an explicit attribute indexer instead"
            )
        categories = json.loads(lines[0].split("\t", 1)[1])
        names = lines[1].split("\t")[1:]
        phonemes, rows = [], []
        for line in lines[2:]:
            cells = line.split("\t")
            phonemes.append(cells[0])
            rows.append([int(v) for v in cells[1:]])
        return cls(phonemes, names, np.array(rows, dtype=np.int64).reshape(len(phonemes), len(names)), categories)


@dataclass
class AllophoneData:
    shared_phone_indexer: ArticulatoryAttributes
    inventories: Dict[str, List[str]] = field(default_factory=dict)  # ISO 639-3 -> phoneme inventory


class PhoneticAttributeIndexer:
    def __init__(
        self,
        full_attributes: ArticulatoryAttributes,
        phoneme_subset: Optional[Sequence[str]] = None,
        composition_features: Optional[Sequence[str]] = None,
        attribute_subset: Optional[Sequence[str]] = None,
        language_allophones: Optional[LanguageAllophoneMappings] = None,
        allophone_data: Optional[AllophoneData] = None,
    ) -> None:
        self._full_attributes = full_attributes
        self._phonemes = _PhonemeIndex(full_attributes.phonemes if phoneme_subset is None else phoneme_subset)
        self._composition_features = list(
            [name for name in full_attributes.feature_names if name != "phoneme"] if composition_features is None else composition_features
        )
        self._attribute_subset = list(full_attributes.feature_names if attribute_subset is None else attribute_subset)
        self._language_allophones = language_allophones
        self._allophone_data = allophone_data

    @property
    def full_attributes(self) -> ArticulatoryAttributes:
        return self._full_attributes

    @property
    def phonemes(self) -> _PhonemeIndex:
        return self._phonemes

    @property
    def composition_features(self) -> List[str]:
        return self._composition_features

    @property
    def language_allophones(self) -> Optional[LanguageAllophoneMappings]:
        return self._language_allophones

    @property
    def allophone_data(self) -> Optional[AllophoneData]:
        return self._allophone_data

    @property
    def feature_names(self) -> List[str]:
        return self._attribute_subset

    def feature_categories(self, name: str) -> List[str]:
        if name == "phoneme":
            return self._phonemes.tolist()
        return self._full_attributes.feature_categories(name)

    def size(self, name: str) -> int:
        return len(self.feature_categories(name))

    def composition_feature_matrix(self, inventory: Sequence[str]) -> Tensor:
        """int64 ``[len(inventory), n_composition_features]`` raw category ids (``phonetic_features.py:808-818``)."""
        return self._full_attributes.subset(list(inventory), self._composition_features).dense_feature_table.long()

    def phoneme_inventory(self, languages: "Sequence[str] | str") -> List[str]:
        if self._allophone_data is None:
            raise ValueError("Allophone inventories can only be accessed if features were extracted from Allophoible")
        codes = [languages] if isinstance(languages, str) else list(languages)
        inventory: List[str] = []
        for code in codes:
            for phoneme in self._allophone_data.inventories[code]:
                if phoneme not in inventory:
                    inventory.append(phoneme)
        return inventory

    # -- checkpoint state (estimator.py:199-249 `phonetic_indexer_state`) --
    def state(self) -> Dict[str, Any]:
        allophones = None
        if self._language_allophones is not None:
            allophones = {
                "allophones": {int(l): {int(p): list(map(int, q)) for p, q in m.items()} for l, m in self._language_allophones.allophones.items()},
                "languages": list(self._language_allophones.languages),
                "shared_phones": list(self._language_allophones.shared_phones),
            }
        return {"phoneme_inventory": self._phonemes.tolist(), "language_allophones": allophones, "table_file": self._full_attributes.to_csv()}

    @classmethod
    def from_state(cls, state: Mapping[str, Any], composition_features: Optional[Sequence[str]] = None) -> "PhoneticAttributeIndexer":
        table_file = state.get("table_file")
        if table_file is None:
            raise ValueError("the checkpoint carries no feature table")
        attributes = ArticulatoryAttributes.from_csv(table_file)
        allophones = state.get("language_allophones")
        mappings = None
        allophone_data = None
        if allophones is not None:
            mappings = LanguageAllophoneMappings(
                {int(l): {int(p): list(map(int, q)) for p, q in m.items()} for l, m in allophones["allophones"].items()},
                list(allophones["languages"]),
                list(allophones["shared_phones"]),
            )
            allophone_data = AllophoneData(attributes.subset(mappings.shared_phones, composition_features))
        return cls(attributes, state["phoneme_inventory"], composition_features, None, mappings, allophone_data)

    @classmethod
    def synthetic(
        cls,
        n_phonemes: int,
        feature_names: Sequence[str],
        n_categories: int = 3,
        seed: int = 1,
        training_inventory: Optional[int] = None,
    ) -> "PhoneticAttributeIndexer":
        """Random categorical table (the real Allophoible CSV is not available offline)."""
        generator = np.random.default_rng(seed)
        names = [name for name in feature_names if name != "phoneme"]
        table = generator.integers(0, n_categories, size=(n_phonemes, len(names)))
        table[:n_categories, :] = np.arange(n_categories)[:, None]  # every category occurs at least once
        phonemes = [f"p{index}" for index in range(n_phonemes)]
        categories = {name: [str(value) for value in range(n_categories)] for name in names}
        attributes = ArticulatoryAttributes(phonemes, names, table, categories)
        subset = phonemes if training_inventory is None else phonemes[:training_inventory]
        return cls(attributes, subset, names, names + ["phoneme"])
