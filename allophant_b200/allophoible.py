"""Ingestion of the Allophoible / PHOIBLE table a reference checkpoint embeds (``phonetic_indexer_state.table_file``).

A pandas-free restatement of the table pipeline of ``allophant/phonetic_features.py``: ``read_allophoible`` (983-999),
``_binarize_contours`` (527-555), ``extract_allophone_inventories`` (1067-1193) with ``_select_largest_inventories`` (1016-1042)
and ``_filter_inventory`` (1045-1064), and ``generate_allophone_data`` (560-598).  It works on a plain row-major table of
strings (``None`` = missing); the orders pandas produces — first-occurrence de-duplication, sorted group keys, descending
size sort, concatenation order — are reproduced because phone indices in a checkpoint's parameters depend on them.

Two pandas behaviours are pinned down explicitly:
  * ``Series.sort_values(ascending=False)`` on the inventory sizes is treated as a STABLE descending sort (equal sizes keep
    the sorted order of their (Source, ISO6393, SpecificDialect) keys; pandas' default quicksort does that for short inputs);
  * the in-place removal of the zero phoneme (1142-1144) is applied (pandas 2 semantics, the version the reference pins).
"""
from __future__ import annotations

import csv
import io
import json
import re
import warnings
from dataclasses import dataclass, field
from importlib import resources
from typing import Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

from .language_codes import LanguageCode

# pandas' default NA strings (read_csv, keep_default_na=True)
_NA_STRINGS = frozenset(
    ["", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"]
)
FEATURE_START = "tone"
_SOURCE_AND_LANGUAGE = ("Source", "ISO6393", "SpecificDialect")
_ZERO_PHONEME = re.compile(r"( ?∅|∅ ?)")


class LanguageMappingWarning(UserWarning):
    """Warns about languages being remapped to a closely related variant (``phonetic_features.py:1004``)."""


@dataclass
class Table:
    """Row-major string table; ``columns`` in file order (``Phoneme`` wherever the file has it)."""

    columns: List[str]
    rows: List[List[Optional[str]]]
    _positions: Dict[str, int] = field(default_factory=dict, repr=False)

    def __post_init__(self) -> None:
        self._positions = {name: index for index, name in enumerate(self.columns)}

    def col(self, name: str) -> int:
        try:
            return self._positions[name]
        except KeyError:
            raise KeyError(f"the feature table has no column {name!r}") from None

    def span(self, first: str, last: str) -> range:
        """Column positions of the label slice ``first:last`` (inclusive, like ``DataFrame.loc``)."""
        return range(self.col(first), self.col(last) + 1)

    @property
    def feature_columns(self) -> List[str]:
        return self.columns[self.col(FEATURE_START) :]


def read_allophoible(text: str) -> Table:
    """``read_allophoible(file, index_column="Phoneme")`` followed by ``reset_index()``: every cell a string or ``None``."""
    reader = csv.reader(io.StringIO(text))
    try:
        header = next(reader)
    except StopIteration:
        raise ValueError("empty feature table") from None
    rows = []
    width = len(header)
    for line in reader:
        if not line:
            continue
        if len(line) != width:
            raise ValueError(f"feature table row with {len(line)} cells, expected {width}")
        rows.append([None if cell in _NA_STRINGS else cell for cell in line])
    table = Table(list(header), rows)
    for required in ("Phoneme", "InventoryID", "ISO6393", "Allophones", "Marginal", "SegmentClass", "Source", "SpecificDialect", FEATURE_START):
        table.col(required)
    inventory = table.col("InventoryID")
    for row in rows:
        int(row[inventory])  # ``astype({"InventoryID": int})`` raises on anything else
    return table


def first_occurrences(table: Table, phonemes: Optional[Iterable[str]] = None) -> List[List[Optional[str]]]:
    """Rows of the first occurrence of every phoneme (``~index.duplicated(keep="first")`` / ``drop_duplicates("Phoneme")``),
    in table order; optionally only those in ``phonemes``."""
    wanted = None if phonemes is None else set(phonemes)
    position = table.col("Phoneme")
    seen = set()
    result = []
    for row in table.rows:
        phoneme = row[position]
        if phoneme in seen or (wanted is not None and phoneme not in wanted):
            continue
        seen.add(phoneme)
        result.append(row)
    return result


def contours(cell: Optional[str]) -> List[str]:
    if cell is None:
        raise ValueError("the feature table has a phone without a value for one of its features")
    return cell.split(",")


def collect_vocabularies(rows: Sequence[Sequence[Optional[str]]], positions: Sequence[int], names: Sequence[str]) -> Dict[str, Dict[str, int]]:
    """``_collect_vocabulary`` per column: values of all contours, sorted, numbered from 0."""
    vocabularies = {}
    for position, name in zip(positions, names):
        values = set()
        for row in rows:
            values.update(contours(row[position]))
        vocabularies[name] = {value: index for index, value in enumerate(sorted(values))}
    return vocabularies


# ---------------------------------------------------------------------------------------------------------------------
# allophone inventories
# ---------------------------------------------------------------------------------------------------------------------
def _sort_key_nan_last(value: Optional[str]) -> Tuple[int, str]:
    return (1, "") if value is None else (0, value)


def select_largest_inventories(
    rows: Sequence[Sequence[Optional[str]]], table: Table, preferred_dialects: Optional[Mapping[str, str]]
) -> List[Tuple[Optional[str], Optional[str], Optional[str]]]:
    """One (Source, ISO6393, SpecificDialect) triple per language: the preferred dialect where one is configured, the
    largest inventory otherwise (``_select_largest_inventories``)."""
    source, iso, dialect = (table.col(name) for name in _SOURCE_AND_LANGUAGE)
    triples = [(row[source], row[iso], row[dialect]) for row in rows]
    if preferred_dialects is not None:
        kept = []
        for language, preferred in preferred_dialects.items():
            kept += [t for t in triples if t[1] == language and t[2] == preferred]
        kept += [t for t in triples if t[1] not in preferred_dialects]
        triples = kept
    sizes: Dict[Tuple[Optional[str], Optional[str], Optional[str]], int] = {}
    for triple in triples:
        sizes[triple] = sizes.get(triple, 0) + 1
    ordered = sorted(sizes, key=lambda t: tuple(_sort_key_nan_last(v) for v in t))  # groupby: sorted keys, NaN last
    ordered.sort(key=lambda t: -sizes[t])  # stable: ties keep the key order
    seen = set()
    result = []
    for triple in ordered:
        if triple[1] in seen:
            continue
        seen.add(triple[1])
        result.append(triple)
    return result


def default_dialects() -> Dict[str, str]:
    with (resources.files(__package__) / "package_data" / "default_dialects.json").open("r", encoding="utf-8") as file:
        return json.load(file)


def extract_allophone_inventories(
    table: Table,
    language_codes: Optional[Sequence[str]] = None,
    remapped_inventories: Optional[Mapping[str, Sequence[str]]] = None,
    prefer_default_dialects: bool = False,
    remove_zero_phoneme: bool = False,
) -> List[List[Optional[str]]]:
    """Rows (copies, in ``table``'s column layout) of the inventories selected for ``language_codes`` — preceded by one row
    (``InventoryID`` 0, no language) for every allophone that is no phoneme of a selected inventory."""
    phoneme, allophones, marginal, iso = table.col("Phoneme"), table.col("Allophones"), table.col("Marginal"), table.col("ISO6393")
    source, dialect = table.col("Source"), table.col("SpecificDialect")
    non_marginal = [row for row in table.rows if row[allophones] is not None and row[marginal] != "TRUE"]
    if language_codes is not None:
        wanted = {LanguageCode.from_str(code).alpha3 for code in language_codes}
        filtered = [row for row in non_marginal if row[iso] in wanted]
    else:
        wanted = None
        filtered = non_marginal
    preferred = default_dialects() if prefer_default_dialects else None
    languages = select_largest_inventories(filtered, table, preferred)

    missing_mappings: Dict[str, str] = {}
    if wanted is not None and len(languages) != len(wanted):
        # languages without an inventory of their own: fall back to a variant of the same macro language
        missing = {LanguageCode.from_str(language, True, True).alpha3_t: language for language in wanted - {t[1] for t in languages}}
        seen_codes = []
        for row in non_marginal:
            if row[iso] not in seen_codes:
                seen_codes.append(row[iso])
        for language in seen_codes:
            macro = LanguageCode.from_str(language, True, True).alpha3_t
            if macro in missing:
                missing_mappings[missing.pop(macro)] = language
            elif language == macro and macro in missing_mappings:
                missing_mappings[missing_mappings[macro]] = language
        if missing:
            raise ValueError(f"Some of the requested languages don't contain allophone data: {sorted(missing.values())}")
        warnings.warn(f"Remapped some languages to a variant within the same macro language: {missing_mappings}", LanguageMappingWarning)
        variants = set(missing_mappings.values())
        languages = languages + select_largest_inventories([row for row in non_marginal if row[iso] in variants], table, preferred)

    selected = set(languages)
    filtered = [list(row) for row in table.rows if (row[source], row[iso], row[dialect]) in selected]
    renamed = {variant: language for language, variant in missing_mappings.items()}
    for row in filtered:
        row[iso] = renamed.get(row[iso], row[iso])

    if remapped_inventories is not None:
        groups: Dict[Optional[str], List[List[Optional[str]]]] = {}
        for row in filtered:
            groups.setdefault(row[iso], []).append(row)
        language_span = table.span("InventoryID", "SpecificDialect")
        filtered = []
        for language in sorted(groups, key=_sort_key_nan_last):
            expected = set(remapped_inventories[language])
            subset = [row for row in groups[language] if row[phoneme] in expected]
            filtered += subset
            remaining = expected - {row[phoneme] for row in subset}
            if not remaining:
                continue
            if not subset:
                raise ValueError(f"none of the phonemes expected for {language!r} occurs in its inventory")
            additions = [list(row) for row in first_occurrences(table, remaining)]
            assert len(additions) == len(remaining), "Inventory mismatch detected"
            for row in additions:
                row[allophones] = row[phoneme]  # only the phoneme itself as its allophone
                for position in language_span:
                    if position != phoneme:
                        row[position] = subset[0][position]
                row[marginal] = None
            filtered += additions

    if remove_zero_phoneme:
        for row in filtered:
            if row[allophones] is not None:
                row[allophones] = _ZERO_PHONEME.sub("", row[allophones])

    phones_in_use = []
    seen = set()
    for row in filtered:
        for phone in [None] if row[allophones] is None else row[allophones].split(" "):
            if phone not in seen:
                seen.add(phone)
                phones_in_use.append(phone)
    missing_phonemes = set(phones_in_use) - {row[phoneme] for row in filtered}
    additional = [list(row) for row in first_occurrences(table, missing_phonemes)]
    without_features = missing_phonemes - {row[phoneme] for row in additional}
    if without_features:
        raise ValueError("Missing pre-computed feature definitions for", len(without_features), "allophones:", without_features)
    inventory = table.col("InventoryID")
    for row in additional:
        row[inventory] = "0"
        for position in table.span("Glottocode", "SpecificDialect"):
            if position != phoneme:
                row[position] = None
        row[source] = None
        row[allophones] = None
    return additional + filtered


@dataclass
class AllophoneInventories:
    """What the reference keeps as the ``AllophoneData.inventories`` frame (``phonetic_features.py:518-521``), reduced to
    the columns that are read: phoneme (the index), ISO 639-3 code, inventory id and the allophone list."""

    phonemes: List[str]
    iso6393: List[Optional[str]]
    inventory_ids: List[int]
    allophones: List[Optional[List[str]]]

    def unique_phonemes(self, codes: Optional[Iterable[str]] = None, database_only: bool = False) -> List[str]:
        wanted = None if codes is None else set(codes)
        seen = set()
        result = []
        for phoneme, iso, inventory in zip(self.phonemes, self.iso6393, self.inventory_ids):
            if (wanted is not None and iso not in wanted) or (database_only and inventory == 0) or phoneme in seen:
                continue
            seen.add(phoneme)
            result.append(phoneme)
        return result

    def language_allophones(self, code: str) -> Dict[str, List[str]]:
        """phoneme -> allophones of one language (``.loc[ISO6393 == code, "Allophones"].str.split(" ").to_dict()``)."""
        return {
            phoneme: allophones
            for phoneme, iso, allophones in zip(self.phonemes, self.iso6393, self.allophones)
            if iso == code and allophones is not None
        }


def allophone_inventories(table: Table, rows: Sequence[Sequence[Optional[str]]]) -> AllophoneInventories:
    phoneme, iso, inventory, allophones = table.col("Phoneme"), table.col("ISO6393"), table.col("InventoryID"), table.col("Allophones")
    return AllophoneInventories(
        [row[phoneme] for row in rows],
        [row[iso] for row in rows],
        [int(row[inventory]) for row in rows],
        [None if row[allophones] is None else row[allophones].split(" ") for row in rows],
    )


def iso6393_inventories(languages: Sequence[str], inventories: Mapping[int, Sequence[str]]) -> Dict[str, List[str]]:
    """``LanguageInventories.iso6393_inventories`` / ``LanguageAllophoneMappings.iso6393_inventories`` (46-52, 103-107)."""
    return {LanguageCode.from_str(language).alpha3: list(inventories[language_id]) for language_id, language in enumerate(languages)}


def binarized(rows: Sequence[Sequence[Optional[str]]], positions: Sequence[int], names: Sequence[str], vocabularies: Mapping[str, Mapping[str, int]]) -> List[List[Tuple[int, ...]]]:
    """``_binarize_vocabulary`` over the given columns: every cell becomes the tuple of category ids of its contour."""
    out = []
    for row in rows:
        cells = []
        for position, name in zip(positions, names):
            vocabulary = vocabularies[name]
            try:
                cells.append(tuple(vocabulary[value] for value in contours(row[position])))
            except KeyError as error:
                raise KeyError(f"feature value {error.args[0]!r} of {name!r} is not part of the vocabulary") from None
        out.append(cells)
    return out
