"""Dependency-free reader of the reference's configuration schema.

Mirrors the field names, defaults and tagged-union keys of ``allophant/config.py``
(``Config`` 924-957, ``Architecture`` 807-849, ``ProjectionConfig`` 679-712,
``ProjectionEntryConfig`` 624-644, ``Wav2Vec2PretrainedConfig`` 760-778,
``EmbeddingCompositionConfig`` 666-676, ``CTCLossConfig`` 547-558) without
marshmallow: ``Config.load`` accepts the nested dict stored in a checkpoint's
``config`` entry or parsed from the TOML file, and ``Config.dump`` writes the same
shape back.  Only the parts the forward/loss path reads are typed; everything else
(optimizer, lr schedule, data, profiling) is carried through as plain dicts.
"""
from __future__ import annotations

import dataclasses
import os
import re
from dataclasses import dataclass, field
from enum import Enum
from re import Pattern
from typing import Any, ClassVar, Dict, List, Mapping, Optional

try:  # Python >= 3.11
    import tomllib as _toml_reader
except ImportError:  # pragma: no cover
    _toml_reader = None


class FeatureSet(Enum):
    PHOIBLE = "phoible"
    PANPHON = "panphon"


class PhonemeLayerType(Enum):
    SHARED = "shared"
    PRIVATE = "private"
    ALLOPHONES = "allophones"


class BatchingMode(Enum):
    FRAMES = "frames"
    UTTERANCES = "utterances"


@dataclass
class CTCLossConfig:
    TYPE: ClassVar[str] = "CTC"
    # Offset for the CTC blank label (config.py:555)
    BLANK_OFFSET: ClassVar[int] = 1

    def get_loss(self):
        from .loss_functions import CTCWrapper

        return CTCWrapper()

    def dump(self) -> Dict[str, Any]:
        return {"type": self.TYPE}


@dataclass
class MultiheadAttentionConfig:
    TYPE: ClassVar[str] = "multi-head-attention"
    num_heads: int = 1
    positional_embeddings: bool = False

    def dump(self) -> Dict[str, Any]:
        return {"type": self.TYPE, "num_heads": self.num_heads, "positional_embeddings": self.positional_embeddings}


def _load_loss(mapping: Optional[Mapping[str, Any]]) -> CTCLossConfig:
    if mapping is None:
        return CTCLossConfig()
    kind = mapping.get("type", CTCLossConfig.TYPE)
    if kind != CTCLossConfig.TYPE:
        raise NotImplementedError(f"Loss type {kind!r} is outside the CTC hot path of allophant_b200")
    return CTCLossConfig()


@dataclass
class ProjectionEntryConfig:
    OUTPUT_DEPENDENCY: ClassVar[str] = "OUTPUT"
    OUTPUT_PATTERN: ClassVar[Pattern] = re.compile(rf"^{OUTPUT_DEPENDENCY}(?:_(\d+))?$")
    PHONEME_LAYER: ClassVar[str] = "phoneme"
    PHONE: ClassVar[str] = "phone"

    name: str
    dependencies: List[str] = field(default_factory=lambda: [ProjectionEntryConfig.OUTPUT_DEPENDENCY])
    time_layer: Optional[MultiheadAttentionConfig] = None
    loss: CTCLossConfig = field(default_factory=CTCLossConfig)

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "ProjectionEntryConfig":
        time_layer = mapping.get("time_layer")
        return cls(
            mapping["name"],
            list(mapping.get("dependencies", [cls.OUTPUT_DEPENDENCY])),
            None
            if time_layer is None
            else MultiheadAttentionConfig(time_layer.get("num_heads", 1), time_layer.get("positional_embeddings", False)),
            _load_loss(mapping.get("loss")),
        )

    def dump(self) -> Dict[str, Any]:
        return {
            "name": self.name,
            "dependencies": list(self.dependencies),
            "time_layer": None if self.time_layer is None else self.time_layer.dump(),
            "loss": self.loss.dump(),
        }


@dataclass
class EmbeddingCompositionConfig:
    embedding_size: int


@dataclass
class ProjectionConfig:
    classes: List[ProjectionEntryConfig]
    feature_set: FeatureSet = FeatureSet.PHOIBLE
    phoneme_layer: PhonemeLayerType = PhonemeLayerType.SHARED
    acoustic_model_dropout: float = 0
    dependency_blanks: bool = True
    allophone_l2_alpha: float = 10
    embedding_composition: Optional[EmbeddingCompositionConfig] = None

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "ProjectionConfig":
        composition = mapping.get("embedding_composition")
        return cls(
            [ProjectionEntryConfig.load(entry) for entry in mapping["classes"]],
            FeatureSet(mapping.get("feature_set", "phoible")),
            PhonemeLayerType(mapping.get("phoneme_layer", "shared")),
            mapping.get("acoustic_model_dropout", 0),
            mapping.get("dependency_blanks", True),
            mapping.get("allophone_l2_alpha", 10),
            None if composition is None else EmbeddingCompositionConfig(int(composition["embedding_size"])),
        )

    def dump(self) -> Dict[str, Any]:
        return {
            "classes": [entry.dump() for entry in self.classes],
            "feature_set": self.feature_set.value,
            "phoneme_layer": self.phoneme_layer.value,
            "acoustic_model_dropout": self.acoustic_model_dropout,
            "dependency_blanks": self.dependency_blanks,
            "allophone_l2_alpha": self.allophone_l2_alpha,
            "embedding_composition": None
            if self.embedding_composition is None
            else {"embedding_size": self.embedding_composition.embedding_size},
        }

    def loss_functions(self) -> Dict[str, Any]:
        return {classifier.name: classifier.loss.get_loss() for classifier in self.classes}


@dataclass
class UnfreezeScheduleConfig:
    feature_encoder_steps: Optional[int] = None
    feature_projection_steps: Optional[int] = None
    encoder_steps: Optional[int] = None


@dataclass
class Wav2Vec2PretrainedConfig:
    TYPE: ClassVar[str] = "wav2vec2-pretrained"

    model_id: str
    freeze_feature_encoder: bool = True
    freeze_feature_projection: bool = False
    freeze_encoder: bool = False
    unfreeze_schedule: Optional[UnfreezeScheduleConfig] = None

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "Wav2Vec2PretrainedConfig":
        schedule = mapping.get("unfreeze_schedule")
        return cls(
            mapping["model_id"],
            mapping.get("freeze_feature_encoder", True),
            mapping.get("freeze_feature_projection", False),
            mapping.get("freeze_encoder", False),
            None if schedule is None else UnfreezeScheduleConfig(**schedule),
        )

    def dump(self) -> Dict[str, Any]:
        return {
            "type": self.TYPE,
            "model_id": self.model_id,
            "freeze_feature_encoder": self.freeze_feature_encoder,
            "freeze_feature_projection": self.freeze_feature_projection,
            "freeze_encoder": self.freeze_encoder,
            "unfreeze_schedule": None if self.unfreeze_schedule is None else dataclasses.asdict(self.unfreeze_schedule),
        }


# ---- from-scratch pre-LN transformer acoustic model (config.py:396-520, 716-735 of the reference) ----
@dataclass
class DropoutConfig:
    TYPE: ClassVar[str] = "dropout"
    rate: float = 0


@dataclass
class LayerNormConfig:
    TYPE: ClassVar[str] = "layer_norm"
    affine: bool = False


@dataclass
class Glu1dConfig:
    TYPE: ClassVar[str] = "glu1d"
    out_channels: int = 0
    kernel: int = 1
    stride: int = 1


@dataclass
class MaxPoolingConfig:
    TYPE: ClassVar[str] = "max_pool"
    size: int = 1


_LAYER_TYPES = {cls.TYPE: cls for cls in (DropoutConfig, LayerNormConfig, Glu1dConfig, MaxPoolingConfig)}


def _dump_keyed(value: Any, key: str = "type") -> Dict[str, Any]:
    return {key: value.TYPE, **dataclasses.asdict(value)}


@dataclass
class SequentialFrontendConfig:
    layers: List[Any] = field(default_factory=list)

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "SequentialFrontendConfig":
        layers = []
        for entry in mapping.get("layers", []):
            kind = entry.get("type")
            if kind not in _LAYER_TYPES:
                raise ValueError(f"Unsupported layer type: {kind!r}")
            layers.append(_LAYER_TYPES[kind](**{k: v for k, v in entry.items() if k != "type"}))
        return cls(layers)

    def dump(self) -> Dict[str, Any]:
        return {"layers": [_dump_keyed(layer) for layer in self.layers]}


@dataclass
class DirectFrontendConfig:
    TYPE: ClassVar[str] = "direct"
    input_dropout: float = 0


@dataclass
class LinearFrontendConfig:
    TYPE: ClassVar[str] = "linear"
    neurons: int = 0
    input_dropout: float = 0


@dataclass
class TransformerConfig:
    TYPE: ClassVar[str] = "transformer"
    feedforward_neurons: int = 0
    heads: int = 1
    activation: str = "relu"
    num_layers: int = 1
    dropout_rate: float = 0
    positional_embeddings: bool = True

    def __post_init__(self) -> None:
        if self.activation not in ("relu", "gelu"):
            raise ValueError(f"activation must be one of relu, gelu; got {self.activation!r}")


@dataclass
class TransformerAcousticModelConfig:
    TYPE: ClassVar[str] = "pre-ln-transformer"

    transformer: TransformerConfig
    frontend: Any = field(default_factory=DirectFrontendConfig)
    sequential_frontend: Optional[SequentialFrontendConfig] = None
    elementwise_affine: bool = False

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "TransformerAcousticModelConfig":
        transformer = TransformerConfig(**{k: v for k, v in mapping["transformer"].items() if k != "type"})
        frontend_map = dict(mapping["frontend"])
        architecture = frontend_map.pop("architecture")
        if architecture == DirectFrontendConfig.TYPE:
            frontend: Any = DirectFrontendConfig(**frontend_map)
        elif architecture == LinearFrontendConfig.TYPE:
            frontend = LinearFrontendConfig(**frontend_map)
        else:
            raise ValueError(f"Unsupported frontend architecture: {architecture!r}")
        sequential = mapping.get("sequential_frontend")
        return cls(transformer, frontend, None if sequential is None else SequentialFrontendConfig.load(sequential), mapping.get("elementwise_affine", False))

    def dump(self) -> Dict[str, Any]:
        return {
            "type": self.TYPE,
            "transformer": dataclasses.asdict(self.transformer),
            "frontend": _dump_keyed(self.frontend, "architecture"),
            "sequential_frontend": None if self.sequential_frontend is None else self.sequential_frontend.dump(),
            "elementwise_affine": self.elementwise_affine,
        }


@dataclass
class UnsupportedAcousticModelConfig:
    """``pre-ln-transformer`` / ``wav2vec2`` acoustic models: parsed, but outside this build's hot path."""

    TYPE: str
    options: Dict[str, Any]

    def dump(self) -> Dict[str, Any]:
        return {"type": self.TYPE, **self.options}


def _load_acoustic_model(mapping: Mapping[str, Any]):
    kind = mapping.get("type")
    if kind == Wav2Vec2PretrainedConfig.TYPE:
        return Wav2Vec2PretrainedConfig.load(mapping)
    if kind == TransformerAcousticModelConfig.TYPE:
        return TransformerAcousticModelConfig.load(mapping)
    if kind in ("wav2vec2",):
        return UnsupportedAcousticModelConfig(kind, {k: v for k, v in mapping.items() if k != "type"})
    raise ValueError(f"Unknown acoustic model type: {kind!r}")


@dataclass
class Architecture:
    batch_size: int
    projection: ProjectionConfig
    acoustic_model: Any
    optimizer: Dict[str, Any] = field(default_factory=dict)
    loss: CTCLossConfig = field(default_factory=CTCLossConfig)
    early_stopping_patience: Optional[int] = None
    batching_mode: BatchingMode = BatchingMode.FRAMES
    language_oversampling_factor: Optional[float] = None
    seed: Optional[int] = None
    maximum_iterations: Optional[int] = None
    clip_norm: Optional[float] = None
    lr_schedule: Optional[Dict[str, Any]] = None
    accumulation_factor: int = 1
    step_size: Optional[int] = None
    mixed_precision: bool = False

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "Architecture":
        return cls(
            mapping["batch_size"],
            ProjectionConfig.load(mapping["projection"]),
            _load_acoustic_model(mapping["acoustic_model"]),
            dict(mapping.get("optimizer", {})),
            _load_loss(mapping.get("loss")),
            mapping.get("early_stopping_patience"),
            BatchingMode(mapping.get("batching_mode", "frames")),
            mapping.get("language_oversampling_factor"),
            mapping.get("seed"),
            mapping.get("maximum_iterations"),
            mapping.get("clip_norm"),
            mapping.get("lr_schedule"),
            mapping.get("accumulation_factor", 1),
            mapping.get("step_size"),
            mapping.get("mixed_precision", False),
        )

    def dump(self) -> Dict[str, Any]:
        return {
            "batch_size": self.batch_size,
            "projection": self.projection.dump(),
            "acoustic_model": self.acoustic_model.dump(),
            "optimizer": dict(self.optimizer),
            "loss": self.loss.dump(),
            "early_stopping_patience": self.early_stopping_patience,
            "batching_mode": self.batching_mode.value,
            "language_oversampling_factor": self.language_oversampling_factor,
            "seed": self.seed,
            "maximum_iterations": self.maximum_iterations,
            "clip_norm": self.clip_norm,
            "lr_schedule": self.lr_schedule,
            "accumulation_factor": self.accumulation_factor,
            "step_size": self.step_size,
            "mixed_precision": self.mixed_precision,
        }


@dataclass
class Config:
    nn: Architecture
    preprocessing: Dict[str, Any] = field(default_factory=dict)
    data: Dict[str, Any] = field(default_factory=dict)
    profiling: Optional[Dict[str, Any]] = None

    @classmethod
    def load(cls, mapping: Mapping[str, Any]) -> "Config":
        return cls(
            Architecture.load(mapping["nn"]),
            dict(mapping.get("preprocessing", {})),
            dict(mapping.get("data", {})),
            mapping.get("profiling"),
        )

    @classmethod
    def from_toml(cls, path: str) -> "Config":
        if _toml_reader is None:  # pragma: no cover
            raise RuntimeError("tomllib is unavailable on this Python version")
        with open(path, "rb") as file:
            return cls.load(_toml_reader.load(file))

    @classmethod
    def default(cls) -> "Config":
        """The Multitask architecture of ``allophant/package_data/default_config.toml``."""
        return cls.from_toml(os.path.join(os.path.dirname(__file__), "package_data", "default_config.toml"))

    def dump(self) -> Dict[str, Any]:
        return {"nn": self.nn.dump(), "preprocessing": dict(self.preprocessing), "data": dict(self.data), "profiling": self.profiling}
