"""Drop-in for ``allophant/network/acoustic_model.py`` (the wav2vec2 branch and the classifier heads).

Same public names, signatures, error behaviour and ``state_dict`` layout as the reference
(``Allophant`` 944-1064, ``Wav2Vec2AcousticModel`` 775-853, ``HierarchicalProjection`` 333-550,
``HierarchicalClassifier`` 271-306, ``EmbeddingCompositionLayer`` 180-234, ``AllophoneMapping``
90-177, ``Predictions`` 908-926); the modules hold fp32 master parameters and hand the
arithmetic to the CUDA engine.  Outputs are time-first ``[T', N, classes]`` like the
reference's; internally they are batch-first contiguous blocks, so the caller's usual
``.transpose(1, 0).contiguous()`` (``run.py:771-772``) is a no-op copy-free view change.
"""
from __future__ import annotations

import math
import os
import typing
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Type, TypeVar

import torch
from torch import LongTensor, Tensor, nn
from torch.nn.parameter import Parameter

from .. import ops
from ..attribute_graph import AttributeGraph, AttributeNode
from ..config import (
    Architecture,
    EmbeddingCompositionConfig,
    PhonemeLayerType,
    ProjectionConfig,
    ProjectionEntryConfig,
    TransformerAcousticModelConfig,
    UnfreezeScheduleConfig,
    Wav2Vec2PretrainedConfig,
)
from ..dataset_processing import Batch
from ..engine import EncoderPlan, PackedEncoder, WorkspaceArena, bucket_samples
from . import frontend
from .transformer import (  # noqa: F401  (drop-in names of acoustic_model.py:34-69, 552-759)
    PreLMTransformerEncoderLayer,
    SinusoidalPositionEmbeddings,
    TransformerAcousticModel,
    TransformerEncoderIntermediate,
)
from .wav2vec2 import Wav2Vec2Weights, encoder_config_for

_PAD_VALUE = torch.finfo(torch.float32).min


def zero_mean_unit_var_norm(features: Tensor, lengths: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """``acoustic_model.py:762-767`` (the mask argument is accepted for signature parity and ignored:
    the kernel derives it from ``lengths``)."""
    return ops.zero_mean_unit_var_norm(features, lengths)


class _ParamHolder(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} holds parameters only; use Allophant.forward (CUDA engine)")


class AllophoneMapping(_ParamHolder):
    """Allophone layer (Li et al., 2020); parameters and buffers as in ``acoustic_model.py:105-136``."""

    _allophone_mask: Tensor

    def __init__(self, shared_phone_count: int, phoneme_count: int, blank_offset: int, language_allophones: Any) -> None:
        super().__init__()
        allophones = language_allophones.allophones
        languages = language_allophones.languages
        self._index_map: Dict[str, int] = {}
        allophone_matrix = torch.zeros(len(languages), shared_phone_count, phoneme_count)
        for dense_index, (language_index, allophone_mapping) in enumerate(allophones.items()):
            matrix = allophone_matrix[dense_index]
            matrix[range(blank_offset), range(blank_offset)] = 1
            self._index_map[languages[language_index]] = dense_index
            for phoneme, phones in allophone_mapping.items():
                matrix[torch.tensor(phones) + blank_offset, phoneme + blank_offset] = 1
        self._allophone_matrices = Parameter(allophone_matrix)
        self.register_buffer("_initialization", allophone_matrix.clone(), persistent=False)
        self.register_buffer("_allophone_mask", ~allophone_matrix.bool(), persistent=False)

    @property
    def index_map(self) -> Dict[str, int]:
        return self._index_map

    def l2_penalty(self) -> Tensor:
        return torch.norm_except_dim(self._allophone_matrices - self._initialization, dim=0).sum()


class EmbeddingCompositionLayer(_ParamHolder):
    """Compositional phone(me) embeddings (Li et al., 2021); state as in ``acoustic_model.py:191-217``."""

    def __init__(self, embedding_size: int, attribute_indexer: Any) -> None:
        super().__init__()
        dense_feature_table = attribute_indexer.dense_feature_table.long().clone()
        num_categories = torch.cat((LongTensor([0]), dense_feature_table.max(0).values)) + 1
        unused_categories = torch.cat(
            (torch.tensor([False]), torch.cat([row.bincount(minlength=int(n)) for row, n in zip(dense_feature_table.T, num_categories[1:])]) == 0)
        )
        category_offsets = num_categories.cumsum(0)[:-1].unsqueeze(0)
        dense_feature_table += category_offsets
        self._attribute_embeddings = nn.EmbeddingBag(int(num_categories.sum()), embedding_size, mode="sum")
        with torch.no_grad():
            self._attribute_embeddings.weight[unused_categories] = 0
        self.register_buffer("_dense_feature_table", dense_feature_table, persistent=False)
        self.register_buffer("_category_offsets", category_offsets, persistent=False)
        self.register_buffer("_scale_factor", torch.tensor(math.sqrt(embedding_size)), persistent=False)
        self.embedding_size = embedding_size


class ProjectingMultiheadAttention(_ParamHolder):
    """Time layer of a classifier head (``acoustic_model.py:237-268``): ``Linear -> LayerNorm (-> + sinusoidal positions) ->
    nn.MultiheadAttention over the frames of each utterance (key-padding mask from the frame counts) -> Dropout``.  The module
    owns the reference's parameters under the reference's names; ``HeadsRuntime`` evaluates it (level GEMM for the input
    projection, ``aph_layernorm_any``, in_proj / out_proj as tcgen05 GEMMs, ``aph_attention_small`` for the heads).  Inference
    only: a training step through a time layer raises."""

    def __init__(self, input_dimensions: int, hidden_dimensions: int, num_heads: int, add_positional_embeddings: bool = False, dropout_rate: float = 0) -> None:
        super().__init__()
        if hidden_dimensions % num_heads != 0:
            raise ValueError("embed_dim must be divisible by num_heads")  # nn.MultiheadAttention's own check
        self.input_projection = nn.Linear(input_dimensions, hidden_dimensions)
        self.positional_embeddings = SinusoidalPositionEmbeddings(hidden_dimensions) if add_positional_embeddings else None
        self.layer_norm = nn.LayerNorm(hidden_dimensions)
        self.attention = nn.MultiheadAttention(hidden_dimensions, num_heads)
        self.dropout = nn.Dropout(dropout_rate)
        self.num_heads = num_heads
        self.hidden_dimensions = hidden_dimensions


class HierarchicalClassifier(_ParamHolder):
    def __init__(
        self,
        time_distributed_layer: "nn.Linear | ProjectingMultiheadAttention",
        composition_layer: Optional[EmbeddingCompositionLayer] = None,
        allophone_layer: Optional[AllophoneMapping] = None,
    ) -> None:
        super().__init__()
        self._lengths_required = not isinstance(time_distributed_layer, nn.Linear)
        self._time_distributed_layer = time_distributed_layer
        self._composition_layer = composition_layer
        self._allophone_layer = allophone_layer


def _process_classifier_dependencies(
    attribute_graph: AttributeGraph, node: AttributeNode, output_features: int, blank_offset: int, dependency_blanks: bool = True
) -> Tuple[int, List[AttributeNode]]:
    layer_input_neurons = 0
    dependencies = []
    for target in node.dependencies:
        attribute_node = attribute_graph.get(target)
        if attribute_node is None:
            attribute_node = AttributeNode(target, output_features)
        elif dependency_blanks:
            attribute_node = attribute_node.with_offset(blank_offset)
        layer_input_neurons += attribute_node.size
        dependencies.append(attribute_node)
    return layer_input_neurons, dependencies


@dataclass
class _HeadSpec:
    """Static description of one classifier, derived once from the graph."""

    name: str
    dependencies: List[AttributeNode]
    in_features: int
    out_features: int  # width of the Linear layer (embedding size when composition is used)
    classes: int  # logits width after composition (training inventory) incl. blank
    level: int = 0


class HierarchicalProjection(nn.Module):
    _OUTPUT_PATTERN = ProjectionEntryConfig.OUTPUT_PATTERN

    def __init__(
        self,
        output_features: int,
        attribute_graph: AttributeGraph,
        blank_offset: int,
        dependency_blanks: bool = True,
        language_allophones: Any = None,
        attribute_indexer: Any = None,
        acoustic_model_dropout_rate: float = 0,
        embedding_composition_config: Optional[EmbeddingCompositionConfig] = None,
    ):
        super().__init__()
        self._acoustic_model_dropout = nn.Dropout(acoustic_model_dropout_rate) if acoustic_model_dropout_rate > 0 else None
        self._uses_allophone_mapping = False

        dependency_names = set(attribute_graph.names())
        if len(dependency_names) < len(attribute_graph):
            raise ValueError("Dependencies contain duplicate keys")
        if any(self._OUTPUT_PATTERN.match(name) for name in dependency_names):
            raise ValueError(f"{ProjectionEntryConfig.OUTPUT_DEPENDENCY!r} is a reserved keyword")

        self._blank_offset = blank_offset
        self._dependency_blanks = dependency_blanks
        self._output_features = output_features
        self._layers = nn.ModuleDict()
        self._ordered_nodes: List[Tuple[str, List[AttributeNode]]] = []
        self._specs: List[_HeadSpec] = []
        required_output_layers = set()

        for node in attribute_graph.sort():
            layer_input_neurons, dependencies = _process_classifier_dependencies(
                attribute_graph, node, output_features, blank_offset, dependency_blanks
            )
            if not dependencies:
                raise ValueError("Each class projection requires a dependency")
            self._ordered_nodes.append((node.name, dependencies))
            required_output_layers.update(d.name for d in dependencies if self._OUTPUT_PATTERN.match(d.name))

            is_phoneme_layer = node.name == ProjectionEntryConfig.PHONEME_LAYER
            if language_allophones is not None and is_phoneme_layer:
                self._uses_allophone_mapping = True
                output_size = len(language_allophones.shared_phones) + blank_offset
            else:
                output_size = node.size + blank_offset

            if is_phoneme_layer and embedding_composition_config is not None:
                projection_output_size = embedding_composition_config.embedding_size
            else:
                projection_output_size = output_size

            if node.time_layer_config is not None:
                time_distributed_layer = ProjectingMultiheadAttention(
                    layer_input_neurons,
                    projection_output_size,
                    node.time_layer_config.num_heads,
                    node.time_layer_config.positional_embeddings,
                    acoustic_model_dropout_rate,
                )
            else:
                time_distributed_layer = nn.Linear(layer_input_neurons, projection_output_size)

            if is_phoneme_layer and embedding_composition_config is not None:
                if attribute_indexer is None:
                    raise ValueError(
                        "Model configuration using attribute embedding composition requires an attribute indexer but got `None`"
                    )
                if not self._uses_allophone_mapping:
                    training_attributes = attribute_indexer.full_attributes.subset(
                        attribute_indexer.phonemes.tolist(), attribute_indexer.composition_features.copy()
                    )
                else:
                    if attribute_indexer.allophone_data is None:
                        raise ValueError(
                            "Model configuration using attribute embedding composition and an allophone layer"
                            " requires allophone data in the attribute indexer with but got `None`"
                        )
                    training_attributes = attribute_indexer.allophone_data.shared_phone_indexer
                if output_size != len(training_attributes) + 1:
                    raise ValueError(
                        f"Length of attributes with blanks ({len(training_attributes) + 1}) need to match"
                        f" the number of phones in the allophone mapping ({output_size})"
                    )
                composition_layer = EmbeddingCompositionLayer(embedding_composition_config.embedding_size, training_attributes)
            else:
                composition_layer = None

            if is_phoneme_layer and self._uses_allophone_mapping:
                allophone_layer = AllophoneMapping(output_size, node.size + blank_offset, blank_offset, language_allophones)
            else:
                allophone_layer = None

            self._layers[node.name] = HierarchicalClassifier(time_distributed_layer, composition_layer, allophone_layer)
            self._specs.append(_HeadSpec(node.name, dependencies, layer_input_neurons, projection_output_size, output_size))

        if not required_output_layers:
            raise ValueError(
                "At least one of the input layers requires {ProjectionEntryConfig.OUTPUT_DEPENDENCY!r} as a dependency"
            )
        self._output_dependencies = sorted(required_output_layers)
        self._assign_levels()

    def _assign_levels(self) -> None:
        level_of: Dict[str, int] = {}
        for spec in self._specs:  # topological order: dependencies come first
            level = 0
            for dependency in spec.dependencies:
                if not self._OUTPUT_PATTERN.match(dependency.name):
                    level = max(level, level_of[dependency.name] + 1)
            spec.level = level
            level_of[spec.name] = level

    def forward(self, *args, **kwargs):
        raise RuntimeError("HierarchicalProjection is evaluated by Allophant.forward (CUDA engine)")

    def l2_penalty(self) -> Optional[Tensor]:
        # The reference's `finally: return None` swallows the value (acoustic_model.py:533-539): always None
        return None

    @property
    def classifier_layers(self) -> nn.ModuleDict:
        return self._layers


class AcousticModel(nn.Module):
    pass


class Wav2Vec2AcousticModel(AcousticModel):
    def __init__(
        self,
        model_id: str,
        sampling_rate: int = 16_000,
        freeze_feature_encoder: bool = True,
        freeze_feature_projection: bool = False,
        freeze_encoder: bool = False,
        load_pretrained_weights: bool = True,
        maximum_encoder_layers: Optional[int] = None,
    ) -> None:
        super().__init__()
        cfg = encoder_config_for(model_id)
        if sampling_rate != cfg.sampling_rate:
            raise ValueError(
                "Audio resampling config and the sampling rate required by Wav2Vec2 do not match. "
                f"Expected {cfg.sampling_rate}kHz, got {sampling_rate}kHz"
            )
        self._model = Wav2Vec2Weights(cfg)
        if load_pretrained_weights:
            self._load_pretrained(model_id)
        # `maximum_encoder_layers` only aliases a ModuleList in the reference and does not truncate
        # compute (acoustic_model.py:801-802); mirrored for state_dict compatibility.
        if maximum_encoder_layers is not None:
            self._model.encoder._layers = self._model.encoder.layers[:maximum_encoder_layers]
        self._model.train(self.training)
        if freeze_feature_encoder:
            self._model.freeze_feature_encoder()
        if freeze_feature_projection:
            for parameter in self._model.feature_projection.parameters():
                parameter.requires_grad = False
        if freeze_encoder:
            for parameter in self._model.encoder.parameters():
                parameter.requires_grad = False

        self._use_attention_mask = cfg.return_attention_mask
        self._normalize = cfg.do_normalize
        self._feature_size = cfg.feature_size
        self._upscale_factor = 1
        self._d_model = cfg.hidden_size
        self._output_size = cfg.hidden_size
        self._sampling_rate = sampling_rate
        self._length_functions = [
            frontend.conv_length(kernel_size, stride, use_padding=False)
            for kernel_size, stride in zip(cfg.conv_kernel, cfg.conv_stride)
        ]
        self._packed = PackedEncoder(self._model)
        self._plans: Dict[Tuple[int, int, int, Tuple[Tuple[int, int], ...]], EncoderPlan] = {}
        # Inference launch lists: inputs are padded up to a multiple of `bucket_frames` frames (0 = exact shapes) and all lists
        # of the model overlay ONE workspace arena, so a ragged stream (MaxFrameBatchSampler) neither allocates nor thrashes a
        # small cache of multi-GB plans.  GroupNorm feature extractors normalise over the padded time axis (HF:302-323): padding
        # further would change their output, so they keep exact shapes.
        self.bucket_frames = int(os.environ.get("APH_BUCKET_FRAMES", "64"))
        self._arena = WorkspaceArena()
        self.plan_builds = 0  # launch lists built so far (bench.py's ragged sweep reports the count after warm-up)

    def _load_pretrained(self, model_id: str) -> None:
        try:
            from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2Model
        except ImportError as error:  # pragma: no cover
            raise RuntimeError("loading pre-trained wav2vec2 weights needs the `transformers` package") from error
        pretrained = Wav2Vec2Model.from_pretrained(model_id)
        self._model.load_state_dict(pretrained.state_dict(), strict=True)

    @property
    def model(self) -> Wav2Vec2Weights:
        return self._model

    @property
    def d_model(self) -> int:
        return self._d_model

    @property
    def output_size(self) -> int:
        return self._output_size

    @property
    def hidden_state_count(self) -> int:
        """Entries of HF's ``hidden_states`` tuple: the encoder input plus one per layer."""
        return self._model.config.num_hidden_layers + 1

    @property
    def feature_size(self) -> int:
        return self._feature_size

    @property
    def upscale_factor(self) -> float:
        return self._upscale_factor

    def downsampled_lengths(self, lengths: Tensor) -> Tensor:
        for convolution_function in self._length_functions:
            lengths = convolution_function(lengths)
        return lengths

    # -- engine access ---------------------------------------------------------------------
    def plan_for(self, n_utt: int, samples: int, ldx: int, hidden_blocks: Dict[int, int], training: bool = False, train_extractor: bool = False) -> EncoderPlan:
        self._packed.ensure()
        key = (n_utt, samples, ldx, tuple(sorted(hidden_blocks.items())), training, train_extractor)
        plan = self._plans.get(key)
        if plan is not None and plan.arena is not None and plan.arena_generation != plan.arena.generation:
            plan = None  # carved from an arena that has been replaced since
            del self._plans[key]
        if plan is None or plan.layout_id != self._packed.layout_id:
            pooled = not training
            limit = 64 if pooled else 4  # pooled launch lists own no memory; training plans keep GBs of activations each
            owned = [k for k, cached in self._plans.items() if (cached.arena is not None) == pooled]
            if plan is None and len(owned) >= limit:
                self._plans.pop(owned[0])
            if plan is not None:
                # same shape, re-allocated operands (the parameters moved): keep the workspaces, rebuild the launch list
                plan.rebind(self._packed)
            else:
                plan = self._build_plan(n_utt, samples, ldx, hidden_blocks, training, train_extractor, self._arena if pooled else None)
            plan.layout_id = self._packed.layout_id
            self._plans[key] = plan
        return plan

    def _build_plan(self, n_utt: int, samples: int, ldx: int, hidden_blocks: Dict[int, int], training: bool, train_extractor: bool,
                    arena: Optional[WorkspaceArena]) -> EncoderPlan:  # fmt: skip
        self.plan_builds += 1
        plan = EncoderPlan(self._packed, n_utt, samples, ldx, hidden_blocks, self._normalize, self._use_attention_mask, training, train_extractor, arena)
        if arena is not None and plan.carver.overflow:
            # the arena is too small for this shape: reserve (with headroom for somewhat larger batches), which voids every list
            # carved from the old buffer, and build again
            needed = plan.carver.offset
            del plan
            for key in [k for k, cached in self._plans.items() if cached.arena is arena]:
                del self._plans[key]
            arena.reserve(int(needed * 1.15), self._packed.device)
            plan = EncoderPlan(self._packed, n_utt, samples, ldx, hidden_blocks, self._normalize, self._use_attention_mask, training, train_extractor, arena)
            assert not plan.carver.overflow
        return plan

    def encode(
        self, batch: Batch, ldx: int, hidden_blocks: Dict[int, int], capture: bool = False, training: bool = False, stochastic: Any = None,
        train_extractor: bool = False,
    ) -> Tuple[EncoderPlan, Tensor]:  # fmt: skip
        audio = batch.audio_features
        if not audio.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the batch to the GPU (`batch.to('cuda')`)")
        if audio.dim() == 3 and audio.shape[-1] == 1:
            audio = audio.squeeze(-1)
        if audio.dim() != 2:
            raise ValueError(f"expected raw audio of shape [batch, samples], got {tuple(audio.shape)}")
        audio = audio.float().contiguous()
        lengths = batch.lengths.to(device=audio.device, dtype=torch.int64).contiguous()
        samples, seq_out = audio.shape[1], None
        if not training and self.bucket_frames > 1 and self._model.config.feat_extract_norm == "layer":
            samples, seq_out = bucket_samples(audio.shape[1], self._model.config, self.bucket_frames)
        plan = self.plan_for(audio.shape[0], samples, ldx, hidden_blocks, training, train_extractor)
        if samples != audio.shape[1]:
            # silence behind the batch's longest utterance: every kernel masks by the per-utterance lengths, frames past the
            # unpadded frame count are dropped from what the caller sees (HeadsRuntime.forward)
            plan.audio_in[:, : audio.shape[1]].copy_(audio)
            plan.audio_in[:, audio.shape[1] :].zero_()
            audio = plan.audio_in
        plan.seq_out = seq_out if seq_out is not None and seq_out != plan.seq else None
        frames = torch.empty(audio.shape[0], device=audio.device, dtype=torch.int64)
        plan.run(audio, lengths, frames, capture, stochastic)
        return plan, frames

    def forward(self, batch: Batch, _predict: bool = False) -> Tuple[List[Tensor], Tensor]:
        """All 25 hidden states, time-first, and the frame counts (``acoustic_model.py:837-853``)."""
        plan, frames = self.encode(batch, self._d_model, {}, capture=True)
        assert plan.captured is not None
        n, seq = plan.n_utt, plan.seq
        shown = plan.seq_out or seq
        return [state.view(n, seq, -1).transpose(0, 1)[:shown] for state in plan.captured], frames


class UnfreezeSchedule:
    def __init__(self, feature_extractor: Optional[int] = None, feature_projection: Optional[int] = None, encoder_steps_remaining: Optional[int] = None):
        self._steps = 0
        self._steps_remaining = [feature_extractor, feature_projection, encoder_steps_remaining]

    def step(self, acoustic_model: AcousticModel):
        if not isinstance(acoustic_model, Wav2Vec2AcousticModel):
            raise ValueError(f"Found an unsupported acoustic module type while updating an unfreeze schedule: {type(acoustic_model)}")
        layers = (acoustic_model._model.feature_extractor, acoustic_model._model.feature_projection, acoustic_model._model.encoder)
        for index, layer in enumerate(layers):
            steps = self._steps_remaining[index]
            if steps is None:
                continue
            steps -= 1
            if steps <= 0:
                steps = None
                for parameter in layer.parameters():
                    parameter.requires_grad = True
            self._steps_remaining[index] = steps

    @classmethod
    def from_config(cls, config: UnfreezeScheduleConfig):
        return UnfreezeSchedule(config.feature_encoder_steps, config.feature_projection_steps, config.encoder_steps)


@dataclass
class Predictions:
    """Named output logit / log-probability batches ``[T', N, classes]`` and the frame count per utterance."""

    outputs: Dict[str, Tensor]
    lengths: Tensor

    def __len__(self) -> int:
        return len(self.lengths)

    def task_count(self) -> int:
        return len(self.outputs)


AllophantCls = TypeVar("AllophantCls", bound="Allophant")


def _highest_specific_output_layer(graph: AttributeGraph) -> Optional[int]:
    output_layer_indices = []
    for node in graph:
        for dependency in node.dependencies:
            output_match = ProjectionEntryConfig.OUTPUT_PATTERN.match(dependency)
            if output_match is not None and (layer_index := output_match.group(1)) is not None:
                output_layer_indices.append(int(layer_index))
    return max(output_layer_indices) + 1 if output_layer_indices else None


class Allophant(nn.Module):
    def __init__(
        self,
        acoustic_model: AcousticModel,
        attribute_graph: AttributeGraph,
        blank_offset: int,
        projection_config: ProjectionConfig,
        attribute_indexer: Any = None,
    ):
        super().__init__()
        self._acoustic_model = acoustic_model
        if attribute_indexer is not None and projection_config.phoneme_layer != PhonemeLayerType.SHARED:
            language_allophones = attribute_indexer.language_allophones
        else:
            language_allophones = None
        self._projection = HierarchicalProjection(
            acoustic_model.output_size,
            attribute_graph,
            blank_offset,
            projection_config.dependency_blanks,
            language_allophones,
            attribute_indexer,
            projection_config.acoustic_model_dropout,
            projection_config.embedding_composition,
        )
        self._classes = list(attribute_graph.names())
        from ..heads import HeadsRuntime  # local import: heads.py imports this module's types

        self._heads = HeadsRuntime(self)

    @property
    def acoustic_model(self) -> AcousticModel:
        return self._acoustic_model

    @property
    def d_model(self) -> int:
        return self._acoustic_model.d_model

    @property
    def feature_size(self) -> int:
        return self._acoustic_model.feature_size

    @property
    def classes(self) -> List[str]:
        return self._classes

    @classmethod
    def from_config(
        cls: Type[AllophantCls],
        architecture: Architecture,
        feature_size: int,
        sample_rate: int,
        attribute_graph: AttributeGraph,
        attribute_indexer: Any = None,
        load_pretrained_weights: bool = True,
    ) -> AllophantCls:
        layer_config = architecture.acoustic_model
        if isinstance(layer_config, Wav2Vec2PretrainedConfig):
            acoustic_model = Wav2Vec2AcousticModel(
                layer_config.model_id,
                sample_rate,
                layer_config.freeze_feature_encoder,
                layer_config.freeze_feature_projection,
                layer_config.freeze_encoder,
                load_pretrained_weights,
                _highest_specific_output_layer(attribute_graph),
            )
        elif getattr(layer_config, "TYPE", None) == "wav2vec2":
            raise NotImplementedError("Training Wav2Vec2 from scratch is not yet implemented")
        elif isinstance(layer_config, TransformerAcousticModelConfig):
            acoustic_model = TransformerAcousticModel.from_config(layer_config, feature_size)
        else:
            raise ValueError(f"Unsupported model type: {type(layer_config)}")
        return cls(acoustic_model, attribute_graph, architecture.loss.BLANK_OFFSET, architecture.projection, attribute_indexer)

    def forward(self, batch: Batch, target_feature_indices: Optional[Tensor] = None, predict: bool = False) -> Predictions:
        return self._heads.forward(batch, target_feature_indices, predict, log_probabilities=False)

    def predict_log_probabilities(self, batch: Batch, target_feature_indices: Optional[Tensor] = None) -> Predictions:
        """Fused ``forward(predict=True)`` + per-head ``log_softmax`` (what ``Estimator.predict`` returns)."""
        return self._heads.forward(batch, target_feature_indices, True, log_probabilities=True)

    def map_allophones(self, phone_logits: Tensor, language_ids: Tensor) -> Tensor:
        return self._heads.map_allophones(phone_logits, language_ids)

    @property
    def upscale_factor(self) -> float:
        return self._acoustic_model.upscale_factor

    @property
    def projection(self) -> HierarchicalProjection:
        return self._projection

    def log_probabilities(self, outputs: Tensor) -> Tensor:
        return ops.log_softmax(outputs)

    def downsampled_lengths(self, lengths: Tensor) -> Tensor:
        return self._acoustic_model.downsampled_lengths(lengths)

    def l2_penalty(self) -> Optional[Tensor]:
        return self._projection.l2_penalty()
