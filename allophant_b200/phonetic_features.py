"""Phoneme/feature tables: the producer of ``composition_feature_matrix`` (input contract of the heads).

The reference derives these tables from the Allophoible CSV with pandas (``allophant/phonetic_features.py:246-971``) and
stores the whole CSV in every checkpoint (``phonetic_indexer_state.table_file``, 647/727-728).  This module keeps the pieces
the model, the decoders and the README usage touch — ``ArticulatoryAttributes`` (``dense_feature_table`` / ``subset`` /
``feature_values``, 246-309), ``PhoneticAttributeIndexer`` (``from_config`` 739-786, ``composition_feature_matrix`` 808-818,
``phoneme_inventory`` 831-857, ``feature_categories``, ``size``, ``attributes``) and the serialisable state (40-44, 111-115) —
over integer contour tables, and rebuilds them from a reference-written state through ``allophant_b200.allophoible`` (a
pandas-free restatement of the CSV pipeline), so ``Estimator.restore`` accepts the reference's checkpoints.
"""
from __future__ import annotations

import io
import json
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import allophoible
from .language_codes import LanguageCode, standardize_to_iso6393


@dataclass
class LanguageAllophoneMappings:
    """language id -> {phoneme index -> [shared phone indices]} (``phonetic_features.py:40-44``)."""

    allophones: Dict[int, Dict[int, List[int]]]
    languages: List[str]
    shared_phones: List[str]

    def iso6393_inventories(self, shared_phoneme_inventory: Sequence[str]) -> Dict[str, List[str]]:
        """ISO 639-3 code -> phonemes of that language (``phonetic_features.py:46-52``)."""
        return allophoible.iso6393_inventories(
            self.languages, {language: [shared_phoneme_inventory[index] for index in phonemes] for language, phonemes in self.allophones.items()}
        )

    @classmethod
    def from_allophone_data(cls, attribute_indexer: "PhoneticAttributeIndexer", languages: List[str]) -> "LanguageAllophoneMappings":
        """Per language: phoneme index -> indices of its allophones among the shared phones (``phonetic_features.py:54-82``)."""
        data = attribute_indexer.allophone_data
        if data is None or not isinstance(data.inventories, allophoible.AllophoneInventories):
            raise ValueError("No allophone data is available in the indexer")
        shared = data.shared_phone_indexer
        allophones = {}
        for language_id, language in enumerate(languages):
            inventory = data.inventories.language_allophones(LanguageCode.from_str(language).alpha3)
            allophones[language_id] = {
                attribute_indexer.phoneme_index(phoneme): [int(index) for index in shared.phoneme_indices(phones)] for phoneme, phones in inventory.items()
            }
        return cls(allophones, list(languages), shared.phonemes.tolist())

    @classmethod
    def from_state(cls, state: Any) -> "LanguageAllophoneMappings":
        """From the dict a checkpoint holds (keys become strings in some serialisations) or an instance."""
        if isinstance(state, cls):
            return state
        get = state.get if isinstance(state, Mapping) else lambda name: getattr(state, name)
        return cls(
            {int(language): {int(phoneme): [int(q) for q in phones] for phoneme, phones in mapping.items()} for language, mapping in get("allophones").items()},
            list(get("languages")),
            list(get("shared_phones")),
        )


@dataclass
class LanguageInventories:
    """language id -> phoneme inventory of the training corpus (``phonetic_features.py:85-113``)."""

    inventories: Dict[int, List[str]]
    languages: List[str]

    def shared_inventory(self) -> List[str]:
        return sorted({phoneme for inventory in self.inventories.values() for phoneme in inventory})

    def iso6393_inventories(self) -> Dict[str, List[str]]:
        return allophoible.iso6393_inventories(self.languages, self.inventories)

    def map_allophones(self, attribute_indexer: "PhoneticAttributeIndexer") -> LanguageAllophoneMappings:
        return LanguageAllophoneMappings(
            {language: {int(p): [int(p)] for p in attribute_indexer.phoneme_indices(inventory)} for language, inventory in self.inventories.items()},
            self.languages,
            attribute_indexer.phonemes.tolist(),
        )


@dataclass
class PhoneticIndexerState:
    phoneme_inventory: List[str]
    language_allophones: Optional[LanguageAllophoneMappings] = None
    table_file: Optional[str] = None


class _PhonemeIndex(list):
    """List of phoneme strings with the ``.tolist()`` the reference's pandas ``Index`` offers."""

    def tolist(self) -> List[str]:
        return list(self)

    to_list = tolist


Contour = Tuple[int, ...]


class ArticulatoryAttributes:
    """Categorical feature table ``[phonemes, features]``; a cell is a CONTOUR of category ids (complex segments carry
    several values per feature), ``dense_feature_table`` holds the first value of each (``phonetic_features.py:246-283``)."""

    def __init__(
        self,
        phonemes: Sequence[str],
        feature_names: Sequence[str],
        table: Optional[np.ndarray],
        feature_categories: Mapping[str, Sequence[str]],
        contours: Optional[Sequence[Sequence[Contour]]] = None,
        reindex_phonemes: bool = False,
    ):
        self._phonemes = _PhonemeIndex(phonemes)
        self._positions = {phoneme: index for index, phoneme in enumerate(self._phonemes)}
        self._feature_names = list(feature_names)
        if contours is None:
            table = np.asarray(table, dtype=np.int64).reshape(len(self._phonemes), len(self._feature_names))
            contours = [[(int(value),) for value in row] for row in table]
        self._contours = [list(row) for row in contours]
        categories = {name: list(feature_categories[name]) for name in self._feature_names}
        if reindex_phonemes and "phoneme" in self._feature_names:
            # phonemes are numbered in subset order (``phonetic_features.py:253-256``)
            column = self._feature_names.index("phoneme")
            for index, row in enumerate(self._contours):
                row[column] = (index,)
            categories["phoneme"] = self._phonemes.tolist()
        self._feature_categories = categories
        self._table = np.array([[cell[0] for cell in row] for row in self._contours], dtype=np.int64).reshape(
            len(self._phonemes), len(self._feature_names)
        )
        # float tensor like the reference's (callers apply `.long()`, acoustic_model.py:194, predictions.py:241)
        self._dense_feature_table = torch.from_numpy(self._table.astype(np.float32))

    @property
    def dense_feature_table(self) -> Tensor:
        return self._dense_feature_table

    @property
    def phonemes(self) -> _PhonemeIndex:
        return self._phonemes

    @property
    def feature_names(self) -> List[str]:
        return self._feature_names

    @property
    def feature_columns(self) -> List[str]:
        return self._feature_names

    def feature_categories(self, name: str) -> List[str]:
        return self._feature_categories[name]

    def feature_category_index(self, name: str) -> int:
        return self._feature_names.index(name)

    def feature_values(self, name: str, feature_indices: Iterable[int]) -> List[str]:
        """Category names of decoded indices (README usage; ``phonetic_features.py:204-206``)."""
        categories = self._feature_categories[name]
        return [categories[int(index)] for index in feature_indices]

    def phoneme_index(self, phoneme: str) -> int:
        try:
            return self._positions[phoneme]
        except KeyError:
            raise KeyError(phoneme) from None

    def phoneme(self, index: "int | np.ndarray") -> "str | List[str]":
        if isinstance(index, (int, np.integer)):
            return self._phonemes[int(index)]
        return [self._phonemes[int(i)] for i in index]

    def phoneme_indices(self, phonemes: Iterable[str]) -> np.ndarray:
        phonemes = list(phonemes)
        missing = [phoneme for phoneme in phonemes if phoneme not in self._positions]
        if missing:
            raise ValueError(f"Missing phonemes: {missing}")
        return np.array([self._positions[phoneme] for phoneme in phonemes], dtype=np.int64)

    def feature_vector(self, phone: "str | int") -> List[np.ndarray]:
        """Contours of one phone, one array per feature (``phonetic_features.py:455-458``)."""
        if isinstance(phone, str):
            phone = self.phoneme_index(phone)
        return [np.array(cell, dtype=np.int64) for cell in self._contours[phone]]

    def simplified_feature_vector(self, phone: "str | int") -> Tensor:
        if isinstance(phone, str):
            phone = self.phoneme_index(phone)
        return self._dense_feature_table[phone]

    def get_named(self, index_or_name: "List[str] | str | int | Tensor | np.ndarray", attribute_index_offset: int = 0) -> Dict[str, Tensor]:
        """Label sequences per feature for a phoneme sequence: contours are concatenated, so a complex segment contributes
        several labels (``phonetic_features.py:181-199``)."""
        if isinstance(index_or_name, list):
            indices = self.phoneme_indices(index_or_name)
        elif isinstance(index_or_name, str):
            indices = np.array([self.phoneme_index(index_or_name)])
        elif isinstance(index_or_name, Tensor):
            indices = index_or_name.numpy()
        else:
            indices = np.atleast_1d(np.asarray(index_or_name))
        if len(indices) == 0:
            return {name: torch.empty(0) for name in self._feature_names}
        return {
            name: torch.tensor([value for index in indices for value in self._contours[int(index)][column]], dtype=torch.int64) + attribute_index_offset
            for column, name in enumerate(self._feature_names)
        }

    def subset(
        self,
        phonemes: Optional[Sequence[str]] = None,
        attribute_subset: Optional[Sequence[str]] = None,
        reindex_phonemes: bool = True,
    ) -> "ArticulatoryAttributes":
        rows = np.arange(len(self._phonemes)) if phonemes is None else self.phoneme_indices(phonemes)
        names = self._feature_names if attribute_subset is None else list(attribute_subset)
        columns = [self._feature_names.index(name) for name in names]
        selected = [self._phonemes[int(r)] for r in rows]
        contours = [[self._contours[int(r)][c] for c in columns] for r in rows]
        return ArticulatoryAttributes(selected, names, None, dict(self._feature_categories), contours, reindex_phonemes)

    def __len__(self) -> int:
        return len(self._phonemes)

    # -- CSV round trip of tables that do not come from an Allophoible file (synthetic inventories) --
    def to_csv(self) -> str:
        buffer = io.StringIO()
        buffer.write("#allophant_b200-feature-table\t" + json.dumps(self._feature_categories, ensure_ascii=False) + "\n")
        buffer.write("phoneme\t" + "\t".join(self._feature_names) + "\n")
        for phoneme, row in zip(self._phonemes, self._contours):
            buffer.write(phoneme + "\t" + "\t".join(",".join(str(int(v)) for v in cell) for cell in row) + "\n")
        return buffer.getvalue()

    @classmethod
    def from_csv(cls, text: str) -> "ArticulatoryAttributes":
        lines = text.splitlines()
        if not lines or not lines[0].startswith("#allophant_b200-feature-table\t"):
            raise ValueError("not a table written by ArticulatoryAttributes.to_csv (Allophoible files go through PhoneticAttributeIndexer.from_allophoible)")
        categories = json.loads(lines[0].split("\t", 1)[1])
        names = lines[1].split("\t")[1:]
        phonemes, rows = [], []
        for line in lines[2:]:
            cells = line.split("\t")
            phonemes.append(cells[0])
            rows.append([tuple(int(v) for v in cell.split(",")) for cell in cells[1:]])
        return cls(phonemes, names, None, categories, rows)


@dataclass
class AllophoneData:
    shared_phone_indexer: ArticulatoryAttributes
    # ISO 639-3 -> phoneme inventory, or the selected database inventories of an Allophoible table
    inventories: "Union[Dict[str, List[str]], allophoible.AllophoneInventories]" = field(default_factory=dict)


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
        self._subset_attributes: Optional[ArticulatoryAttributes] = None
        self._full_phoneme_subset_attributes: Optional[ArticulatoryAttributes] = None
        self._table_file: Optional[str] = None
        self._positions = {phoneme: index for index, phoneme in enumerate(self._phonemes)}

    # -- construction from an Allophoible table ----------------------------------------------------
    @classmethod
    def from_allophoible(
        cls,
        table_file: str,
        attribute_subset: Optional[Sequence[str]] = None,
        phoneme_subset: Optional[Sequence[str]] = None,
        language_inventories: "Union[LanguageInventories, LanguageAllophoneMappings, Sequence[str], None]" = None,
        allophones_from_allophoible: bool = False,
    ) -> "PhoneticAttributeIndexer":
        """The PHOIBLE branch of the reference's constructor (``phonetic_features.py:601-725``): full feature table with
        contours, the classifier subset, the selected language inventories and the shared-phone table of the allophone layer."""
        if not isinstance(table_file, str):
            table_file = table_file.read()
        table = allophoible.read_allophoible(table_file)

        # ---- allophone inventories of the requested languages (generate_allophone_data, 560-598)
        if isinstance(language_inventories, LanguageInventories):
            languages: Optional[Sequence[str]] = language_inventories.languages
            remapped = language_inventories.iso6393_inventories()
        elif isinstance(language_inventories, LanguageAllophoneMappings):
            languages = language_inventories.languages
            if phoneme_subset is None:
                raise ValueError("allophone inventories can only be restored from LanguageAllophoneMappings if a correct phoneme_subset is provided")
            remapped = language_inventories.iso6393_inventories(phoneme_subset)
        elif language_inventories is None:
            languages, remapped = None, None
        else:
            languages, remapped = list(language_inventories), None
        allophone_rows = allophoible.extract_allophone_inventories(table, languages, remapped, prefer_default_dialects=True, remove_zero_phoneme=True)
        inventories = allophoible.allophone_inventories(table, allophone_rows)
        if phoneme_subset is None:
            phoneme_subset = inventories.unique_phonemes(database_only=True)

        # ---- full table: first row of every phone, contours binarised against the sorted vocabulary of each column (621-661)
        unique_rows = allophoible.first_occurrences(table)
        feature_names = table.feature_columns
        positions = [table.col(name) for name in feature_names]
        vocabularies = allophoible.collect_vocabularies(unique_rows, positions, feature_names)
        phoneme_column = table.col("Phoneme")
        full_phonemes = [row[phoneme_column] for row in unique_rows]
        contours = allophoible.binarized(unique_rows, positions, feature_names, vocabularies)
        for index, row in enumerate(contours):
            row.append((index,))
        categories: Dict[str, List[str]] = {name: list(vocabulary) for name, vocabulary in vocabularies.items()}
        categories["phoneme"] = list(full_phonemes)
        full = ArticulatoryAttributes(full_phonemes, [*feature_names, "phoneme"], None, categories, contours)

        attribute_subset = None if attribute_subset is None else list(attribute_subset)
        composition_features = [name for name in feature_names[1:] if name != "phoneme"]  # everything behind "tone" (684-701)
        indexer = cls(full, phoneme_subset, composition_features, attribute_subset)
        indexer._subset_attributes = full.subset(phoneme_subset, attribute_subset)
        indexer._phonemes = indexer._subset_attributes.phonemes
        indexer._positions = {phoneme: index for index, phoneme in enumerate(indexer._phonemes)}
        indexer._attribute_subset = list(indexer._subset_attributes.feature_names)
        full_subset = attribute_subset if attribute_subset is None or "phoneme" in attribute_subset else [*attribute_subset, "phoneme"]
        indexer._full_phoneme_subset_attributes = full.subset(attribute_subset=full_subset)
        indexer._table_file = table_file

        # ---- shared phones of the allophone layer: unique phones of the selected inventories, features behind "tone" (703-713)
        shared_names = feature_names[1:]
        shared_positions = positions[1:]
        seen = set()
        shared_rows = []
        for row in allophone_rows:
            if row[phoneme_column] not in seen:
                seen.add(row[phoneme_column])
                shared_rows.append(row)
        shared = ArticulatoryAttributes(
            [row[phoneme_column] for row in shared_rows],
            shared_names,
            None,
            {name: categories[name] for name in shared_names},
            allophoible.binarized(shared_rows, shared_positions, shared_names, vocabularies),
        )
        indexer._allophone_data = AllophoneData(shared, inventories)

        if isinstance(language_inventories, LanguageAllophoneMappings):
            indexer._language_allophones = language_inventories
        elif isinstance(language_inventories, LanguageInventories):
            if allophones_from_allophoible:
                indexer._language_allophones = LanguageAllophoneMappings.from_allophone_data(indexer, language_inventories.languages)
            else:
                indexer._language_allophones = language_inventories.map_allophones(indexer)
        return indexer

    @classmethod
    def from_config(
        cls,
        config: Any,
        attribute_table_file: Optional[str] = None,
        language_inventories: Optional[LanguageInventories] = None,
        state_dict: "Union[PhoneticIndexerState, Mapping[str, Any], None]" = None,
    ) -> "PhoneticAttributeIndexer":
        """``phonetic_features.py:739-786``: the classifier names of the configuration select the attribute subset; a
        checkpoint's ``phonetic_indexer_state`` supplies the table, the training inventory and the allophone mappings."""
        from .config import FeatureSet, PhonemeLayerType, ProjectionEntryConfig

        projection = config.nn.projection
        if projection.feature_set != FeatureSet.PHOIBLE:
            raise NotImplementedError("only the PHOIBLE / Allophoible feature set is supported (panphon is not available)")
        entries: Dict[str, None] = {}
        for entry in projection.classes:
            entries[entry.name] = None
            entries.update((attribute, None) for attribute in entry.dependencies)
        entries.pop(ProjectionEntryConfig.OUTPUT_DEPENDENCY, None)
        for attribute in list(entries):
            if ProjectionEntryConfig.OUTPUT_PATTERN.match(attribute):
                del entries[attribute]

        if isinstance(state_dict, Mapping):
            state_dict = PhoneticIndexerState(
                list(state_dict["phoneme_inventory"]), state_dict.get("language_allophones"), state_dict.get("table_file")
            )
        mappings: "Union[LanguageInventories, LanguageAllophoneMappings, None]"
        if state_dict is not None and state_dict.language_allophones is not None:
            mappings = LanguageAllophoneMappings.from_state(state_dict.language_allophones)
            phoneme_subset: Optional[List[str]] = list(state_dict.phoneme_inventory)
            attribute_table_file = state_dict.table_file
        elif language_inventories is not None:
            mappings = language_inventories
            phoneme_subset = sorted(language_inventories.shared_inventory())
        else:
            mappings = phoneme_subset = None
        if attribute_table_file is None:
            raise ValueError("an Allophoible feature table is required (the database file is not bundled with this package)")
        return cls.from_allophoible(
            attribute_table_file, list(entries), phoneme_subset, mappings, projection.phoneme_layer == PhonemeLayerType.ALLOPHONES
        )

    # -- accessors ---------------------------------------------------------------------------------
    @property
    def full_attributes(self) -> ArticulatoryAttributes:
        return self._full_attributes

    @property
    def attributes(self) -> ArticulatoryAttributes:
        """The table of the training phonemes over the classifier attributes (``phonetic_features.py:796-798``)."""
        if self._subset_attributes is None:
            names = [name for name in self._attribute_subset if name in self._full_attributes.feature_names]
            self._subset_attributes = self._full_attributes.subset(self._phonemes.tolist(), names)
        return self._subset_attributes

    @property
    def full_subset_attributes(self) -> ArticulatoryAttributes:
        if self._full_phoneme_subset_attributes is None:
            names = [name for name in self._attribute_subset if name in self._full_attributes.feature_names]
            self._full_phoneme_subset_attributes = self._full_attributes.subset(attribute_subset=names)
        return self._full_phoneme_subset_attributes

    @property
    def phonemes(self) -> _PhonemeIndex:
        return self._phonemes

    def phoneme_index(self, phoneme: str) -> int:
        return self._positions[phoneme]

    def phoneme_indices(self, phonemes: Iterable[str]) -> np.ndarray:
        phonemes = list(phonemes)
        missing = [phoneme for phoneme in phonemes if phoneme not in self._positions]
        if missing:
            raise ValueError(f"Missing phonemes: {missing}")
        return np.array([self._positions[phoneme] for phoneme in phonemes], dtype=np.int64)

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
        if name == "phone" and self._language_allophones is not None:
            return list(self._language_allophones.shared_phones)
        return self._full_attributes.feature_categories(name)

    def feature_values(self, name: str, feature_indices: Iterable[int]) -> List[str]:
        categories = self.feature_categories(name)
        return [categories[int(index)] for index in feature_indices]

    def size(self, name: Optional[str] = None) -> int:
        if name is None:
            return sum(len(self.feature_categories(feature)) for feature in self._attribute_subset)
        return len(self.feature_categories(name))

    def composition_feature_matrix(self, inventory: Sequence[str]) -> Tensor:
        """int64 ``[len(inventory), n_composition_features]`` raw category ids (``phonetic_features.py:808-818``)."""
        return self._full_attributes.subset(list(inventory), self._composition_features).dense_feature_table.long()

    def phoneme_inventory(self, languages: "Sequence[str] | str") -> List[str]:
        """Union of the database inventories of the given languages (``phonetic_features.py:831-857``)."""
        if self._allophone_data is None:
            raise ValueError("Allophone inventories can only be accessed if features were extracted from Allophoible")
        codes = [languages] if isinstance(languages, str) else list(languages)
        inventories = self._allophone_data.inventories
        if isinstance(inventories, allophoible.AllophoneInventories):
            return inventories.unique_phonemes({standardize_to_iso6393(code) for code in codes})
        inventory: List[str] = []
        for code in codes:
            for phoneme in inventories[code]:
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
        table_file = self._table_file if self._table_file is not None else self._full_attributes.to_csv()
        return {"phoneme_inventory": self._phonemes.tolist(), "language_allophones": allophones, "table_file": table_file}

    @classmethod
    def from_state(
        cls, state: Mapping[str, Any], composition_features: Optional[Sequence[str]] = None, config: Any = None
    ) -> "PhoneticAttributeIndexer":
        """Rebuilds the indexer of a checkpoint: a reference-written state (embedded Allophoible CSV) goes through
        ``from_config`` like ``Estimator.restore`` does in the reference (``estimator.py:1110-1112``); states written by
        this package for synthetic tables carry their own compact table format."""
        table_file = state.get("table_file") if isinstance(state, Mapping) else state.table_file
        if table_file is None:
            raise ValueError("the checkpoint carries no feature table")
        if not table_file.startswith("#allophant_b200-feature-table\t"):
            if config is None:
                raise ValueError("restoring an Allophoible-backed indexer needs the checkpoint's configuration (classifier names)")
            return cls.from_config(config, state_dict=state)
        attributes = ArticulatoryAttributes.from_csv(table_file)
        allophones = state.get("language_allophones")
        mappings = None
        allophone_data = None
        if allophones is not None:
            mappings = LanguageAllophoneMappings.from_state(allophones)
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
