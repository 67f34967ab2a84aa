"""TEST INFRASTRUCTURE — CPU restatement of the reference's acoustic-model forward/loss path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker or the timed CPU baseline.  The product
(``allophant_b200``) never imports it.

What is restated (plain torch, fp32, CPU) and from where:
  * ``zero_mean_unit_var_norm``            allophant/network/acoustic_model.py:762-767
  * ``mask_sequence``                      allophant/utils.py:45-76
  * ``Wav2Vec2AcousticModel.forward``      allophant/network/acoustic_model.py:837-853 (+ 832-835, frontend.py:192-203)
  * ``HierarchicalProjection.forward``     allophant/network/acoustic_model.py:471-524 (+ 284-306, 309-330)
  * ``EmbeddingCompositionLayer``          allophant/network/acoustic_model.py:180-234
  * ``AllophoneMapping``                   allophant/network/acoustic_model.py:75-87, 142-167
  * ``Estimator.predict``                  allophant/estimator.py:1035-1046
  * ``CTCWrapper``                         allophant/loss_functions.py:19-27
  * ``GreedyCTCDecoder``                   allophant/predictions.py:189-207
  * step loss arithmetic + backward        allophant/estimator.py:708-738 (``OracleModel.training_step``)

The encoder arithmetic itself lives in a third-party dependency of the reference,
``transformers`` (pinned 4.41.2 in the reference's pyproject.toml:23; 5.5.0 is what this image
has): ``Wav2Vec2Model`` is called exactly like the reference calls it.  CTC and log_softmax are
torch's (the reference calls ``nn.CTCLoss`` / ``functional.log_softmax``).

PARITY PIN: the reference has no tests or golden vectors for this path, so
the pin is generated here: ``oracle/make_golden.py`` runs the UNMODIFIED reference classes from
/root/reference (through ``oracle/reference_shim.py``) and this restatement on the same seeds, checks
they agree, and freezes the reference's outputs in ``tests/golden/``; ``tests/test_oracle_golden.py``
re-checks the restatement against those files wherever the tests run.
"""
from __future__ import annotations

import contextlib
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor, nn
from torch.nn import functional

OUTPUT_PATTERN = re.compile(r"^OUTPUT(?:_(\d+))?$")  # config.py:637
BLANK_OFFSET = 1  # config.py:555
_PAD_VALUE = torch.finfo(torch.float32).min  # acoustic_model.py:72

XLSR_300M = dict(
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

PHOIBLE_FEATURES = [
    "stress", "syllabic", "short", "long", "consonantal", "sonorant", "continuant", "delayedRelease", "approximant",
    "tap", "trill", "nasal", "lateral", "labial", "round", "labiodental", "coronal", "anterior", "distributed",
    "strident", "dorsal", "high", "low", "front", "back", "tense", "retractedTongueRoot", "advancedTongueRoot",
    "periodicGlottalSource", "epilaryngealSource", "spreadGlottis", "constrictedGlottis", "fortis",
    "raisedLarynxEjective", "loweredLarynxImplosive", "click",
]  # fmt: skip


# --------------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------------
def mask_sequence(lengths: Tensor, max_length: Optional[int] = None, inverse: bool = False) -> Tensor:
    """utils.py:45-76 (batch-first form)."""
    if max_length is None:
        max_length = int(lengths.max())
    positions = torch.arange(0, max_length, device=lengths.device).unsqueeze(0)
    return positions >= lengths.unsqueeze(1) if inverse else positions < lengths.unsqueeze(1)


def zero_mean_unit_var_norm(features: Tensor, lengths: Tensor, mask: Tensor) -> Tensor:
    """acoustic_model.py:762-767."""
    means = (features.sum(1) / lengths).unsqueeze(1)
    deviations = (features - means) * mask
    variances = (deviations**2).sum(1) / lengths
    return ((features - means) / (variances.unsqueeze(1) + 1e-7).sqrt()) * mask


def conv_lengths(lengths: Tensor, kernels: Sequence[int], strides: Sequence[int]) -> Tensor:
    """frontend.py:192-203 (use_padding=False) folded over the layers, acoustic_model.py:832-835."""
    for kernel, stride in zip(kernels, strides):
        lengths = torch.div(lengths - kernel, stride, rounding_mode="floor") + 1
    return lengths


class CTCHypothesis(NamedTuple):
    tokens: Tensor
    words: List[str]
    score: Tensor
    timesteps: Tensor


def greedy_ctc_decode(log_emissions: Tensor, lengths: Tensor, blank_index: int = 0) -> List[List[CTCHypothesis]]:
    """predictions.py:194-207; ``log_emissions`` is batch-first ``[N, T', classes]``."""
    batch_max = torch.max(log_emissions, dim=-1)
    outputs = []
    for i, indices in enumerate(batch_max.indices):
        length = lengths[i]
        indices = indices[:length]
        decoded, sizes = torch.unique_consecutive(indices, return_counts=True)
        non_blanks = decoded != blank_index
        timesteps = (sizes.cumsum(0) - sizes + 1)[non_blanks]
        outputs.append([CTCHypothesis(decoded[non_blanks], [], batch_max.values[i, :length].sum(), timesteps)])
    return outputs


def ctc_wrapper(logits: Tensor, labels: Tensor, predicted_lengths: Tensor, label_lengths: Tensor) -> Tensor:
    """loss_functions.py:19-27: nn.CTCLoss(reduction="sum", zero_infinity=True) on log_softmax(logits)."""
    return functional.ctc_loss(
        functional.log_softmax(logits, -1), labels, predicted_lengths, label_lengths, blank=0, reduction="sum", zero_infinity=True
    )


def multiply_allophone_matrix(phone_logits: Tensor, matrix: Tensor, mask: Tensor) -> Tensor:
    """acoustic_model.py:75-87."""
    return (phone_logits * matrix.unsqueeze(0)).masked_fill_(mask.unsqueeze(0), _PAD_VALUE).max(1).values


# --------------------------------------------------------------------------------------------------
# model description
# --------------------------------------------------------------------------------------------------
@dataclass
class ClassSpec:
    name: str
    size: int  # number of categories (without blank)
    dependencies: List[str] = field(default_factory=lambda: ["OUTPUT"])


@dataclass
class OracleSpec:
    """Everything needed to rebuild a reference-shaped model deterministically from seeds."""

    classes: List[ClassSpec]
    feature_table: np.ndarray  # int [P_train, F] raw category ids of the training inventory (composition)
    embedding_size: Optional[int] = 640
    dependency_blanks: bool = True
    encoder_overrides: Dict[str, object] = field(default_factory=dict)
    weight_seed: int = 2
    # allophone layer (phoneme_layer != shared): language -> {phoneme index -> [phone indices]}
    allophones: Optional[Dict[int, Dict[int, List[int]]]] = None
    n_phones: int = 0


def multitask_spec(
    n_categories: int = 3,
    n_train_phonemes: int = 60,
    table_seed: int = 1,
    encoder_overrides: Optional[Dict[str, object]] = None,
    hierarchical: bool = False,
    weight_seed: int = 2,
) -> OracleSpec:
    """The reference's Multitask (default_config.toml) or Hierarchical architecture over a synthetic table."""
    rng = np.random.default_rng(table_seed)
    table = rng.integers(0, n_categories, size=(n_train_phonemes, len(PHOIBLE_FEATURES)))
    table[:n_categories, :] = np.arange(n_categories)[:, None]
    classes = [ClassSpec(name, n_categories) for name in PHOIBLE_FEATURES]
    phoneme_dependencies = ["OUTPUT", *PHOIBLE_FEATURES] if hierarchical else ["OUTPUT"]
    classes.append(ClassSpec("phoneme", n_train_phonemes, phoneme_dependencies))
    return OracleSpec(classes, table, 640, True, dict(encoder_overrides or {}), weight_seed)


def topological_order(classes: Sequence[ClassSpec]) -> List[ClassSpec]:
    """attribute_graph.py:124-199 for acyclic graphs: DFS post-order from node 0, edges in listed order."""
    index = {spec.name: i for i, spec in enumerate(classes)}
    edges = [[index[d] for d in spec.dependencies if not OUTPUT_PATTERN.match(d)] for spec in classes]
    seen, order = set(), []

    def visit(node: int, chain: Tuple[int, ...]) -> None:
        if node in chain:
            raise ValueError("Dependency cycle detected")
        if node in seen:
            return
        seen.add(node)
        for target in edges[node]:
            visit(target, chain + (node,))
        order.append(node)

    for node in range(len(classes)):
        visit(node, ())
    return [classes[i] for i in order]


class OracleModel:
    """Reference-shaped model on CPU: HF ``Wav2Vec2Model`` + restated heads, seeded like the reference builds it."""

    def __init__(self, spec: OracleSpec) -> None:
        from transformers import Wav2Vec2Config
        from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2Model

        self.spec = spec
        config = dict(XLSR_300M)
        config.update(spec.encoder_overrides)
        torch.manual_seed(spec.weight_seed)
        # acoustic_model.py:798: Wav2Vec2Model(Wav2Vec2Config.from_pretrained(model_id)) — created first
        self.encoder = Wav2Vec2Model(Wav2Vec2Config(**config))
        self.encoder.eval()
        self.config = self.encoder.config
        self.ordered = topological_order(spec.classes)
        sizes = {c.name: c.size for c in spec.classes}
        hidden = self.config.hidden_size
        self.linears: Dict[str, nn.Linear] = {}
        self.composition: Optional[nn.EmbeddingBag] = None
        self.allophone_matrices: Optional[Tensor] = None
        # acoustic_model.py:368-461: layers are created in topological order
        for node in self.ordered:
            in_features = 0
            for dependency in node.dependencies:
                if OUTPUT_PATTERN.match(dependency):
                    in_features += hidden
                else:
                    in_features += sizes[dependency] + (BLANK_OFFSET if spec.dependency_blanks else 0)
            is_phoneme = node.name == "phoneme"
            if is_phoneme and spec.allophones is not None:
                output_size = spec.n_phones + BLANK_OFFSET
            else:
                output_size = node.size + BLANK_OFFSET
            if is_phoneme and spec.embedding_size is not None:
                projection_size = spec.embedding_size
            else:
                projection_size = output_size
            self.linears[node.name] = nn.Linear(in_features, projection_size)
            if is_phoneme and spec.embedding_size is not None:
                # EmbeddingCompositionLayer.__init__, acoustic_model.py:191-217
                table = torch.from_numpy(np.asarray(spec.feature_table)).long()
                num_categories = torch.cat((torch.LongTensor([0]), table.max(0).values)) + 1
                unused = torch.cat((torch.tensor([False]), torch.cat([row.bincount() for row in table.T]) == 0))
                self.category_offsets = num_categories.cumsum(0)[:-1].unsqueeze(0)
                self.dense_feature_table = table + self.category_offsets
                self.composition = nn.EmbeddingBag(int(num_categories.sum()), spec.embedding_size, mode="sum")
                with torch.no_grad():
                    self.composition.weight[unused] = 0
                self.scale_factor = torch.tensor(math.sqrt(spec.embedding_size))
            if is_phoneme and spec.allophones is not None:
                # AllophoneMapping.__init__, acoustic_model.py:105-136
                matrix = torch.zeros(len(spec.allophones), output_size, node.size + BLANK_OFFSET)
                for dense_index, (_, mapping) in enumerate(spec.allophones.items()):
                    matrix[dense_index][range(BLANK_OFFSET), range(BLANK_OFFSET)] = 1
                    for phoneme, phones in mapping.items():
                        matrix[dense_index][torch.tensor(phones) + BLANK_OFFSET, phoneme + BLANK_OFFSET] = 1
                self.allophone_matrices = matrix.clone()
                self.allophone_mask = ~matrix.bool()

    # ------------------------------------------------------------------ state
    def state_dict(self) -> Dict[str, Tensor]:
        """Same keys as the reference ``Allophant.state_dict()``."""
        state = {f"_acoustic_model._model.{k}": v.detach().clone() for k, v in self.encoder.state_dict().items()}
        for name, linear in self.linears.items():
            state[f"_projection._layers.{name}._time_distributed_layer.weight"] = linear.weight.detach().clone()
            state[f"_projection._layers.{name}._time_distributed_layer.bias"] = linear.bias.detach().clone()
        if self.composition is not None:
            state["_projection._layers.phoneme._composition_layer._attribute_embeddings.weight"] = self.composition.weight.detach().clone()
        if self.allophone_matrices is not None:
            state["_projection._layers.phoneme._allophone_layer._allophone_matrices"] = self.allophone_matrices.clone()
        return state

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def encode(self, audio: Tensor, lengths: Tensor) -> Tuple[List[Tensor], Tensor]:
        return self._encode(audio, lengths)

    @torch.no_grad()
    def compose(self, inputs: Tensor, target_feature_indices: Optional[Tensor]) -> Tensor:
        return self._compose(inputs, target_feature_indices)

    @torch.no_grad()
    def project(
        self,
        hidden_states: List[Tensor],
        language_ids: Optional[Tensor] = None,
        target_feature_indices: Optional[Tensor] = None,
        predict: bool = True,
    ) -> Dict[str, Tensor]:
        return self._project(hidden_states, language_ids, target_feature_indices, predict)

    def _encode(self, audio: Tensor, lengths: Tensor, regularisation: Optional[Dict[str, object]] = None) -> Tuple[List[Tensor], Tensor]:
        """Wav2Vec2AcousticModel.forward, acoustic_model.py:837-853 (do_normalize / return_attention_mask = True).

        ``regularisation``: the train()-mode stochastic ops of the HF encoder with EXPLICIT masks instead of torch's RNG
        (see ``explicit_regularisation``), so that a CUDA run with counter-based masks can be compared exactly."""
        mask = mask_sequence(lengths)
        frames = conv_lengths(lengths, self.config.conv_kernel, self.config.conv_stride)
        normed = zero_mean_unit_var_norm(audio, lengths, mask)
        if regularisation is None:
            hidden_states = self.encoder(normed, mask.long(), output_hidden_states=True).hidden_states
        else:
            with self.explicit_regularisation(regularisation):
                hidden_states = self.encoder(
                    normed, mask.long(), output_hidden_states=True, mask_time_indices=regularisation.get("spec")
                ).hidden_states
            if "classifier_input" in regularisation:  # acoustic_model.py:486-488 (batch-first masks by hidden-state index)
                hidden_states = list(hidden_states)
                for index, factor in regularisation["classifier_input"].items():
                    hidden_states[index] = hidden_states[index] * factor
        return [h.transpose(0, 1) for h in hidden_states], frames

    @contextlib.contextmanager
    def explicit_regularisation(self, masks: Dict[str, object]):
        """Runs the eval()-mode HF encoder with the train()-mode ops applied through EXPLICIT multiplicative masks:
        ``feature_projection`` / ``encoder_input`` / ``attention_output.<l>`` / ``feed_forward_output.<l>``: fp32
        ``[N, T', H]`` (keep / (1-p) or 0), ``activation.<l>``: ``[N, T', FF]`` (``activation_dropout``), ``attention.<l>``: ``[N, heads, T', T']``, ``skip``: LayerDrop decisions
        (HF modeling_wav2vec2.py:431-433, 766, 774-786, 458, 643, 572); ``spec`` (bool ``[N, T']``) is passed to the model as
        ``mask_time_indices``."""
        from transformers.models.wav2vec2 import modeling_wav2vec2 as hf

        handles = []

        def scale_output(module: nn.Module, key: str) -> None:
            if key in masks:
                handles.append(module.register_forward_hook(lambda _m, _i, out, key=key: out * masks[key]))

        scale_output(self.encoder.feature_projection.dropout, "feature_projection")
        scale_output(self.encoder.encoder.dropout, "encoder_input")
        skip = masks.get("skip") or []
        for index, layer in enumerate(self.encoder.encoder.layers):
            scale_output(layer.dropout, f"attention_output.{index}")
            scale_output(layer.feed_forward.output_dropout, f"feed_forward_output.{index}")
            scale_output(layer.feed_forward.intermediate_dropout, f"activation.{index}")
            layer.attention._oracle_index = index
            if index < len(skip) and skip[index]:
                handles.append(layer.register_forward_hook(lambda _m, args, _out: (args[0],)))
        original, implementation = hf.eager_attention_forward, self.encoder.config._attn_implementation
        mask_hidden_states = self.encoder._mask_hidden_states
        if "spec_feature" in masks:  # HF zeroes the feature-axis spans after the time mask (bool [N, H])

            def masked(hidden_states, *args, **kwargs):
                hidden_states = mask_hidden_states(hidden_states, *args, **kwargs)
                return hidden_states * (~masks["spec_feature"])[:, None, :].to(hidden_states.dtype)

            self.encoder._mask_hidden_states = masked

        def attention(module, query, key, value, attention_mask, scaling=None, dropout=0.0, **kwargs):
            weights = torch.matmul(query, key.transpose(2, 3)) * (query.size(-1) ** -0.5 if scaling is None else scaling)
            if attention_mask is not None:
                weights = weights + attention_mask
            weights = nn.functional.softmax(weights, dim=-1)
            factor = masks.get(f"attention.{module._oracle_index}")
            if factor is not None:
                weights = weights * factor
            return torch.matmul(weights, value).transpose(1, 2).contiguous(), weights

        hf.eager_attention_forward = attention
        self.encoder.config._attn_implementation = "eager"
        try:
            yield
        finally:
            hf.eager_attention_forward = original
            self.encoder.config._attn_implementation = implementation
            if "spec_feature" in masks:
                del self.encoder._mask_hidden_states  # back to the class's method
            for handle in handles:
                handle.remove()

    def _compose(self, inputs: Tensor, target_feature_indices: Optional[Tensor]) -> Tensor:
        """EmbeddingCompositionLayer.forward, acoustic_model.py:219-234."""
        assert self.composition is not None
        if target_feature_indices is None:
            indices = self.dense_feature_table
        else:
            indices = target_feature_indices + self.category_offsets
        composed = torch.cat((self.composition(torch.zeros(1, 1, dtype=indices.dtype)), self.composition(indices))).T
        return (inputs @ composed) / self.scale_factor

    def _project(
        self,
        hidden_states: List[Tensor],
        language_ids: Optional[Tensor] = None,
        target_feature_indices: Optional[Tensor] = None,
        predict: bool = True,
    ) -> Dict[str, Tensor]:
        """HierarchicalProjection.forward, acoustic_model.py:471-524 (eval mode: no dropout)."""
        outputs = {f"OUTPUT_{i}": h for i, h in enumerate(hidden_states)}
        outputs["OUTPUT"] = hidden_states[-1]
        skip = 0 if self.spec.dependency_blanks else BLANK_OFFSET
        results: Dict[str, Tensor] = {}
        for node in self.ordered:
            if len(node.dependencies) == 1 and OUTPUT_PATTERN.match(node.dependencies[0]):
                inputs = outputs[node.dependencies[0]]
            else:
                inputs = torch.cat(
                    [
                        outputs[d] if OUTPUT_PATTERN.match(d) else torch.softmax(outputs[d][..., skip:], -1)
                        for d in node.dependencies
                    ],
                    -1,
                )
            out = self.linears[node.name](inputs)
            if node.name == "phoneme" and self.composition is not None:
                out = self._compose(out, target_feature_indices)
            if node.name == "phoneme" and self.allophone_matrices is not None:
                if predict:
                    produced = {"phone": out, "phoneme": out}  # acoustic_model.py:164-166
                else:
                    columns = []
                    for index, language_id in enumerate(map(int, language_ids)):  # acoustic_model.py:142-159
                        columns.append(
                            multiply_allophone_matrix(
                                out[:, index].unsqueeze(-1), self.allophone_matrices[language_id], self.allophone_mask[language_id]
                            )
                        )
                    produced = {"phoneme": torch.stack(columns, 1)}
                results.update(produced)
                outputs.update(produced)
            else:
                results[node.name] = out
                outputs[node.name] = out
        return results

    @torch.no_grad()
    def predict(
        self,
        audio: Tensor,
        lengths: Tensor,
        language_ids: Optional[Tensor] = None,
        target_feature_indices: Optional[Tensor] = None,
        log_probabilities: bool = True,
    ) -> Tuple[Dict[str, Tensor], Tensor]:
        """Estimator.predict, estimator.py:1035-1046 → (name -> [T', N, classes], frames [N])."""
        hidden_states, frames = self.encode(audio, lengths)
        logits = self.project(hidden_states, language_ids, target_feature_indices, predict=True)
        if log_probabilities:
            return {name: functional.log_softmax(value, -1) for name, value in logits.items()}, frames
        return logits, frames


    # ------------------------------------------------------------------ training step
    def trainable_parameters(self, freeze_feature_encoder: bool = True) -> Dict[str, Tensor]:
        """Leaf tensors under their ``Allophant.state_dict()`` names; the convolutional feature extractor is frozen
        like the reference's default (``default_config.toml:40``, ``acoustic_model.py:806-807``)."""
        named: Dict[str, Tensor] = {}
        for key, parameter in self.encoder.named_parameters():
            frozen = freeze_feature_encoder and key.startswith("feature_extractor.")
            parameter.requires_grad_(not frozen)
            if not frozen:
                named[f"_acoustic_model._model.{key}"] = parameter
        for name, linear in self.linears.items():
            named[f"_projection._layers.{name}._time_distributed_layer.weight"] = linear.weight
            named[f"_projection._layers.{name}._time_distributed_layer.bias"] = linear.bias
        if self.composition is not None:
            named["_projection._layers.phoneme._composition_layer._attribute_embeddings.weight"] = self.composition.weight
        if self.allophone_matrices is not None:
            self.allophone_matrices.requires_grad_(True)
            named["_projection._layers.phoneme._allophone_layer._allophone_matrices"] = self.allophone_matrices
        return named

    def training_step(
        self,
        audio: Tensor,
        lengths: Tensor,
        labels: Dict[str, Tensor],
        label_lengths: Dict[str, Tensor],
        language_ids: Optional[Tensor] = None,
        target_feature_indices: Optional[Tensor] = None,
        regularisation: Optional[Dict[str, object]] = None,
        freeze_feature_encoder: bool = True,
    ) -> Tuple[Tensor, Dict[str, float], Dict[str, Tensor]]:
        """``estimator.py:708-738`` in eval()-mode arithmetic (or with explicit train()-mode masks): ``model(batch)`` (predict=False), per-head
        ``CTCWrapper``, ``loss = sum_heads ctc / sum_heads sum_utt label_length``, ``backward()``.
        Returns (loss, per-head CTC sums, gradients by state_dict name)."""
        named = self.trainable_parameters(freeze_feature_encoder)
        for parameter in named.values():
            parameter.grad = None
        with torch.enable_grad():
            hidden_states, frames = self._encode(audio, lengths, regularisation)
            outputs = self._project(hidden_states, language_ids, target_feature_indices, predict=False)
            outputs.pop("phone", None)  # estimator.py:717-718
            total = torch.tensor(0, dtype=torch.float32)
            per_head: Dict[str, float] = {}
            normaliser = 0
            for name, output in outputs.items():
                head_loss = ctc_wrapper(output, labels[name], frames, label_lengths[name])
                per_head[name] = float(head_loss)
                normaliser += int(label_lengths[name].sum())
                total = total + head_loss
            loss = total / normaliser
            loss.backward()
        gradients = {name: parameter.grad.detach().clone() for name, parameter in named.items() if parameter.grad is not None}
        return loss.detach(), per_head, gradients


def state_checksum(state: Dict[str, Tensor]) -> Dict[str, float]:
    """Order-independent fingerprint used to check that seeded weights were regenerated identically."""
    total = 0.0
    absolute = 0.0
    for key in sorted(state):
        value = state[key].double()
        total += float(value.sum())
        absolute += float(value.abs().sum())
    return {"sum": total, "abs_sum": absolute, "tensors": float(len(state))}


def synthetic_audio(n_utt: int, samples: int, seed: int = 0) -> Tensor:
    """``0.1 * randn(N, T)`` under ``torch.manual_seed(seed)``."""
    generator = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(n_utt, samples, generator=generator)


def synthetic_labels(frames: Tensor, n_classes: int, seed: int, fraction: float = 0.25) -> Tuple[Tensor, Tensor]:
    """Random label sequences in ``[1, n_classes)`` of length ``floor(fraction * frames)``."""
    generator = torch.Generator().manual_seed(seed)
    lengths = torch.div(frames.double() * fraction, 1, rounding_mode="floor").long()
    longest = int(lengths.max())
    labels = torch.zeros(len(frames), max(longest, 1), dtype=torch.long)
    for index, length in enumerate(lengths.tolist()):
        labels[index, :length] = torch.randint(1, n_classes, (length,), generator=generator)
    return labels, lengths


def training_labels(
    spec: OracleSpec, head_classes: Dict[str, int], frames: Tensor, language_ids: Optional[Tensor], seed: int = 100
) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Per-head CTC labels for a training step: random sequences of length
    ``floor(0.25 * frames)``.  With an allophone layer the phoneme labels of an utterance are drawn from its own
    language's inventory: absent phonemes are fully masked (``acoustic_model.py:72, 81-84``), which makes the CTC
    loss infinite and ``zero_infinity`` would silently zero the whole utterance."""
    labels: Dict[str, Tensor] = {}
    lengths: Dict[str, Tensor] = {}
    for index, (name, classes) in enumerate(head_classes.items()):
        head_labels, head_lengths = synthetic_labels(frames, classes, seed=seed + index)
        if name == "phoneme" and spec.allophones is not None:
            assert language_ids is not None
            generator = torch.Generator().manual_seed(seed + 1000)
            languages = list(spec.allophones)
            for row, language in enumerate(language_ids.tolist()):
                inventory = torch.tensor(sorted(spec.allophones[languages[language]]), dtype=torch.long) + BLANK_OFFSET
                length = int(head_lengths[row])
                picks = torch.randint(0, len(inventory), (length,), generator=generator)
                head_labels[row, :length] = inventory[picks]
        labels[name], lengths[name] = head_labels, head_lengths
    return labels, lengths


def gradient_summary(gradient: Tensor, samples: int = 32) -> Dict[str, object]:
    """Compact fingerprint of one gradient tensor for the golden fixtures: norm, sum and evenly spaced samples."""
    flat = gradient.detach().double().flatten()
    index = torch.linspace(0, flat.numel() - 1, min(samples, flat.numel())).long()
    return dict(norm=float(flat.norm()), sum=float(flat.sum()), index=index, values=flat[index].float(), shape=tuple(gradient.shape))
