"""TEST INFRASTRUCTURE — import shim that runs the UNMODIFIED reference from /root/reference.

Nothing in ``allophant_b200`` (the product) imports this module.  It exists so
that ``oracle/make_golden.py`` can execute the reference's own
``Allophant`` / ``Estimator.predict`` / ``CTCWrapper`` / ``GreedyCTCDecoder``
(``allophant/network/acoustic_model.py``, ``allophant/estimator.py:1035-1046``,
``allophant/loss_functions.py:19-27``, ``allophant/predictions.py:189-207``) in
this container and freeze their outputs as golden vectors under
``tests/golden/``.  It cannot travel to the GPU box (``/root/reference`` does
not exist there); the plain-torch restatement in ``oracle/restatement.py`` does,
and is pinned against the vectors produced here.

Why a shim is needed: the reference imports marshmallow,
mashumaro, panphon, langcodes, mutagen, stanza, epitran, phonemizer, zarr, the
Rust extension ``allophant.phonemes`` and ``torchaudio.models.decoder``
(flashlight-text) at module import time.  None of them is installed and none is
touched by the forward/loss arithmetic; they are replaced with permissive dummy
modules.  The two Hugging Face hub calls made unconditionally by
``Wav2Vec2AcousticModel.__init__`` (``acoustic_model.py:787,798``) are patched to
return the XLS-R-300M constants (hub unreachable offline).
"""
from __future__ import annotations

import dataclasses
import importlib
import importlib.abc
import importlib.machinery
import importlib.metadata
import sys
import types
from typing import Any, List, NamedTuple

REFERENCE_ROOT = "/root/reference"

_STUBBED_ROOTS = (
    "marshmallow",
    "marshmallow_dataclass",
    "marshmallow_enum",
    "marshmallow_oneofschema",
    "mashumaro",
    "panphon",
    "langcodes",
    "mutagen",
    "stanza",
    "epitran",
    "phonemizer",
    "zarr",
    "flashlight",
)
_STUBBED_EXACT = ("allophant.phonemes",)


class _DummyMeta(type):
    def __getattr__(cls, name: str) -> Any:
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy

    def __getitem__(cls, item: Any) -> Any:
        return cls

    def __or__(cls, other: Any) -> Any:
        return cls

    def __ror__(cls, other: Any) -> Any:
        return cls


class _Dummy(metaclass=_DummyMeta):
    """Callable, subscriptable, subclassable stand-in for anything the shimmed packages export."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        pass

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        # Used as a decorator (``@validates(...)``, ``@post_load``) it must return the function unchanged
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]
        return _Dummy()

    def __getattr__(self, name: str) -> Any:
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, item: Any) -> Any:
        return _Dummy()

    def __iter__(self):
        return iter(())

    def __init_subclass__(cls, **kwargs: Any) -> None:
        pass


class _DummySchema(_Dummy):
    def load(self, data: Any, *args: Any, **kwargs: Any) -> Any:
        raise RuntimeError("reference_shim: marshmallow Schema.load is not available (marshmallow is stubbed)")

    def dump(self, data: Any, *args: Any, **kwargs: Any) -> Any:
        raise RuntimeError("reference_shim: marshmallow Schema.dump is not available (marshmallow is stubbed)")


def _md_dataclass(_cls: Any = None, **kwargs: Any) -> Any:
    """``marshmallow_dataclass.dataclass`` → ``dataclasses.dataclass`` + a dummy ``.Schema``."""
    kwargs.pop("base_schema", None)

    def wrap(cls: Any) -> Any:
        cls = dataclasses.dataclass(cls, **kwargs)
        cls.Schema = _DummySchema
        return cls

    return wrap if _cls is None else wrap(_cls)


def _md_add_schema(_cls: Any = None, **kwargs: Any) -> Any:
    def wrap(cls: Any) -> Any:
        cls.Schema = _DummySchema
        return cls

    return wrap if _cls is None else wrap(_cls)


class _StubModule(types.ModuleType):
    def __getattr__(self, name: str) -> Any:
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Dummy


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname: str, path: Any, target: Any = None):
        root = fullname.split(".")[0]
        if root in _STUBBED_ROOTS or fullname in _STUBBED_EXACT:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        module = _StubModule(spec.name)
        module.__path__ = []  # behaves as a package so submodule imports resolve through this finder
        if spec.name == "marshmallow_dataclass":
            module.dataclass = _md_dataclass
            module.add_schema = _md_add_schema
            module.class_schema = lambda cls, *a, **k: _DummySchema
        if spec.name == "marshmallow":
            module.Schema = _DummySchema
        return module

    def exec_module(self, module) -> None:
        pass


class CTCHypothesis(NamedTuple):
    """Same fields as ``torchaudio.models.decoder.CTCHypothesis`` (which needs flashlight-text to import)."""

    tokens: Any
    words: List[str]
    score: float
    timesteps: Any


# XLS-R-300M constants (public facebook/wav2vec2-xls-r-300m config.json / preprocessor_config.json;
# SURVEY.md §8c).  Not verifiable offline.
XLSR_300M_CONFIG = dict(
    hidden_size=1024,
    num_hidden_layers=24,
    num_attention_heads=16,
    intermediate_size=4096,
    hidden_act="gelu",
    hidden_dropout=0.1,
    activation_dropout=0.0,
    attention_dropout=0.1,
    feat_proj_dropout=0.1,
    final_dropout=0.0,
    layerdrop=0.1,
    layer_norm_eps=1e-5,
    feat_extract_norm="layer",
    feat_extract_activation="gelu",
    conv_dim=(512, 512, 512, 512, 512, 512, 512),
    conv_stride=(5, 2, 2, 2, 2, 2, 2),
    conv_kernel=(10, 3, 3, 3, 3, 2, 2),
    conv_bias=True,
    num_conv_pos_embeddings=128,
    num_conv_pos_embedding_groups=16,
    do_stable_layer_norm=True,
    mask_time_prob=0.075,
    mask_time_length=10,
    mask_feature_prob=0.0,
    vocab_size=32,
)
XLSR_300M_FEATURE_EXTRACTOR = dict(
    feature_size=1, sampling_rate=16000, padding_value=0, do_normalize=True, return_attention_mask=True
)

_INSTALLED = False
_ENCODER_OVERRIDES: dict = {}


def set_encoder_overrides(**overrides: Any) -> None:
    """Shrinks the encoder (e.g. ``num_hidden_layers=2``) for fast golden cases.  Test-only."""
    _ENCODER_OVERRIDES.clear()
    _ENCODER_OVERRIDES.update(overrides)


def install() -> None:
    """Installs the stubs, patches the hub calls and puts /root/reference on sys.path (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    sys.meta_path.insert(0, _StubFinder())

    # torchaudio.models.decoder raises without flashlight-text: pre-seed a replacement module
    import torchaudio.models  # noqa: F401

    decoder = types.ModuleType("torchaudio.models.decoder")
    decoder.CTCHypothesis = CTCHypothesis
    decoder.CTCDecoder = _Dummy
    decoder.ctc_decoder = _Dummy()
    sys.modules["torchaudio.models.decoder"] = decoder
    sys.modules["torchaudio.models"].decoder = decoder

    # pandas 3 dropped this name (allophant/phonetic_features.py:19)
    import pandas.io.parsers.readers as readers

    if not hasattr(readers, "ReadCsvBuffer"):
        readers.ReadCsvBuffer = object

    # allophant is not installed as a distribution (predictions.py:46, evaluation.py:62)
    original_version = importlib.metadata.version

    def version(name: str) -> str:
        if name == "allophant":
            return "1.0.0"
        return original_version(name)

    importlib.metadata.version = version

    # Hub calls (acoustic_model.py:787,798)
    from transformers import Wav2Vec2Config
    from transformers.models.wav2vec2.feature_extraction_wav2vec2 import Wav2Vec2FeatureExtractor

    def get_feature_extractor_dict(model_id: str, **kwargs: Any):
        return dict(XLSR_300M_FEATURE_EXTRACTOR), {}

    def config_from_pretrained(model_id: str, **kwargs: Any):
        merged = dict(XLSR_300M_CONFIG)
        merged.update(_ENCODER_OVERRIDES)
        return Wav2Vec2Config(**merged)

    Wav2Vec2FeatureExtractor.get_feature_extractor_dict = staticmethod(get_feature_extractor_dict)
    Wav2Vec2Config.from_pretrained = staticmethod(config_from_pretrained)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


def reference_modules():
    """Returns the unmodified reference modules on the hot path."""
    install()
    acoustic_model = importlib.import_module("allophant.network.acoustic_model")
    estimator = importlib.import_module("allophant.estimator")
    loss_functions = importlib.import_module("allophant.loss_functions")
    predictions = importlib.import_module("allophant.predictions")
    config = importlib.import_module("allophant.config")
    attribute_graph = importlib.import_module("allophant.attribute_graph")
    phonetic_features = importlib.import_module("allophant.phonetic_features")
    batching = importlib.import_module("allophant.batching")
    return types.SimpleNamespace(
        acoustic_model=acoustic_model,
        estimator=estimator,
        loss_functions=loss_functions,
        predictions=predictions,
        config=config,
        attribute_graph=attribute_graph,
        phonetic_features=phonetic_features,
        batching=batching,
    )
