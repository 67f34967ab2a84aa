"""Language code standardisation (the part of ``allophant/language_codes.py:8-58`` the feature tables use).

The reference delegates to the ``langcodes`` package (absent here): ``LanguageCode.from_str(code).alpha3`` is the ISO 639-3
(terminology form) code of a BCP 47 tag, and ``from_str(code, True, True)`` first replaces an individual language by its macro
language (``cmn`` -> ``zh``).  This module carries the ISO 639-1 -> 639-3 table, the 20 bibliographic -> terminology pairs and
the individual -> macro language pairs of the languages PHOIBLE / Common Voice inventories use; unknown three-letter codes
pass through unchanged (ISO 639-3 codes are their own alpha-3 form).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

# ISO 639-1 -> ISO 639-3 (terminology form; macro languages keep their macro code, as langcodes' to_alpha3 does)
_ALPHA2 = {
    "aa": "aar", "ab": "abk", "ae": "ave", "af": "afr", "ak": "aka", "am": "amh", "an": "arg", "ar": "ara", "as": "asm", "av": "ava",
    "ay": "aym", "az": "aze", "ba": "bak", "be": "bel", "bg": "bul", "bi": "bis", "bm": "bam", "bn": "ben", "bo": "bod", "br": "bre",
    "bs": "bos", "ca": "cat", "ce": "che", "ch": "cha", "co": "cos", "cr": "cre", "cs": "ces", "cu": "chu", "cv": "chv", "cy": "cym",
    "da": "dan", "de": "deu", "dv": "div", "dz": "dzo", "ee": "ewe", "el": "ell", "en": "eng", "eo": "epo", "es": "spa", "et": "est",
    "eu": "eus", "fa": "fas", "ff": "ful", "fi": "fin", "fj": "fij", "fo": "fao", "fr": "fra", "fy": "fry", "ga": "gle", "gd": "gla",
    "gl": "glg", "gn": "grn", "gu": "guj", "gv": "glv", "ha": "hau", "he": "heb", "hi": "hin", "ho": "hmo", "hr": "hrv", "ht": "hat",
    "hu": "hun", "hy": "hye", "hz": "her", "ia": "ina", "id": "ind", "ie": "ile", "ig": "ibo", "ii": "iii", "ik": "ipk", "io": "ido",
    "is": "isl", "it": "ita", "iu": "iku", "ja": "jpn", "jv": "jav", "ka": "kat", "kg": "kon", "ki": "kik", "kj": "kua", "kk": "kaz",
    "kl": "kal", "km": "khm", "kn": "kan", "ko": "kor", "kr": "kau", "ks": "kas", "ku": "kur", "kv": "kom", "kw": "cor", "ky": "kir",
    "la": "lat", "lb": "ltz", "lg": "lug", "li": "lim", "ln": "lin", "lo": "lao", "lt": "lit", "lu": "lub", "lv": "lav", "mg": "mlg",
    "mh": "mah", "mi": "mri", "mk": "mkd", "ml": "mal", "mn": "mon", "mr": "mar", "ms": "msa", "mt": "mlt", "my": "mya", "na": "nau",
    "nb": "nob", "nd": "nde", "ne": "nep", "ng": "ndo", "nl": "nld", "nn": "nno", "no": "nor", "nr": "nbl", "nv": "nav", "ny": "nya",
    "oc": "oci", "oj": "oji", "om": "orm", "or": "ori", "os": "oss", "pa": "pan", "pi": "pli", "pl": "pol", "ps": "pus", "pt": "por",
    "qu": "que", "rm": "roh", "rn": "run", "ro": "ron", "ru": "rus", "rw": "kin", "sa": "san", "sc": "srd", "sd": "snd", "se": "sme",
    "sg": "sag", "si": "sin", "sk": "slk", "sl": "slv", "sm": "smo", "sn": "sna", "so": "som", "sq": "sqi", "sr": "srp", "ss": "ssw",
    "st": "sot", "su": "sun", "sv": "swe", "sw": "swa", "ta": "tam", "te": "tel", "tg": "tgk", "th": "tha", "ti": "tir", "tk": "tuk",
    "tl": "tgl", "tn": "tsn", "to": "ton", "tr": "tur", "ts": "tso", "tt": "tat", "tw": "twi", "ty": "tah", "ug": "uig", "uk": "ukr",
    "ur": "urd", "uz": "uzb", "ve": "ven", "vi": "vie", "vo": "vol", "wa": "wln", "wo": "wol", "xh": "xho", "yi": "yid", "yo": "yor",
    "za": "zha", "zh": "zho", "zu": "zul",
}  # fmt: skip
_ALPHA3_TO_ALPHA2 = {three: two for two, three in _ALPHA2.items()}

# ISO 639-2 bibliographic -> terminology
_BIBLIOGRAPHIC = {
    "alb": "sqi", "arm": "hye", "baq": "eus", "bur": "mya", "chi": "zho", "cze": "ces", "dut": "nld", "fre": "fra", "geo": "kat",
    "ger": "deu", "gre": "ell", "ice": "isl", "mac": "mkd", "mao": "mri", "may": "msa", "per": "fas", "rum": "ron", "slo": "slk",
    "tib": "bod", "wel": "cym",
}  # fmt: skip

# individual language -> macro language (ISO 639-3 macrolanguage mappings of the codes PHOIBLE inventories carry)
_MACRO = {
    "cmn": "zho", "yue": "zho", "wuu": "zho", "hak": "zho", "nan": "zho", "gan": "zho", "hsn": "zho", "cdo": "zho",
    "arb": "ara", "arz": "ara", "apc": "ara", "ary": "ara", "acm": "ara", "afb": "ara", "ajp": "ara", "aeb": "ara", "arq": "ara",
    "ekk": "est", "vro": "est", "lvs": "lav", "ltg": "lav", "pes": "fas", "prs": "fas", "zsm": "msa", "zlm": "msa", "ind": "msa",
    "swh": "swa", "swc": "swa", "uzn": "uzb", "uzs": "uzb", "khk": "mon", "mvf": "mon", "npi": "nep", "dty": "nep", "ory": "ori",
    "spv": "ori", "plt": "mlg", "azj": "aze", "azb": "aze", "als": "sqi", "aln": "sqi", "aae": "sqi", "aat": "sqi", "ydd": "yid",
    "yih": "yid", "quz": "que", "quy": "que", "qub": "que", "gug": "grn", "kmr": "kur", "ckb": "kur", "sdh": "kur", "pbu": "pus",
    "pst": "pus", "pbt": "pus", "nob": "nor", "nno": "nor", "hbs": "hbs", "srp": "hbs", "hrv": "hbs", "bos": "hbs", "gaz": "orm",
    "hae": "orm", "fuv": "ful", "fuf": "ful", "ffm": "ful", "knc": "kau", "kng": "kon", "ike": "iku", "ikt": "iku", "ojg": "oji",
    "crk": "cre", "ayr": "aym", "ayc": "aym", "kpv": "kom", "koi": "kom", "sme": "sme", "bho": "bho", "mai": "mai", "gom": "kok",
    "knn": "kok", "dgo": "doi", "mwr": "mwr", "raj": "raj", "zyb": "zha", "zch": "zha", "hmn": "hmn", "bjn": "msa", "min": "msa",
    "twi": "aka", "fat": "aka",
}  # fmt: skip


@dataclass
class LanguageCode:
    """Same fields as the reference's dataclass (``language_codes.py:8-13``)."""

    language: str
    alpha3_t: str
    alpha3_b: str
    variant: Optional[str]

    @classmethod
    def from_str(cls, language_code: str, standardize: bool = False, macro: bool = False) -> "LanguageCode":
        if macro and not standardize:
            raise ValueError("Retrieving the macro language requires standardization")
        parts = language_code.replace("_", "-").split("-")
        primary = parts[0].lower()
        if not primary.isalpha() or len(primary) not in (2, 3):
            raise ValueError(f"{language_code!r} does not contain a valid language code")
        if len(primary) == 2:
            alpha3 = _ALPHA2.get(primary)
            if alpha3 is None:
                raise ValueError(f"{language_code!r} does not contain a valid language code")
        else:
            alpha3 = _BIBLIOGRAPHIC.get(primary, primary)
        if macro:
            alpha3 = _MACRO.get(alpha3, alpha3)
        language = _ALPHA3_TO_ALPHA2.get(alpha3, alpha3)
        variant = "-".join(parts[1:]) if len(parts) > 1 else None
        terminology = alpha3
        bibliographic = next((b for b, t in _BIBLIOGRAPHIC.items() if t == alpha3), alpha3)
        return cls(language, terminology, bibliographic, variant)

    @property
    def alpha3(self) -> str:
        return self.alpha3_t

    def __str__(self) -> str:
        return self.language if self.variant is None else f"{self.language}-{self.variant}"


def standardize_to_iso6393(language_code: str) -> str:
    """``language_codes.py:56-57``"""
    return LanguageCode.from_str(language_code, True).alpha3
