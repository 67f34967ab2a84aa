"""``Estimator`` facade (drop-in for the inference surface of ``allophant/estimator.py:931-1126``).

``from_config`` / ``restore`` / ``predict`` / ``save`` keep the reference's signatures and the
checkpoint dictionary layout (``estimator.py:199-249``) so a checkpoint written by either side
loads on the other, as far as the pieces in scope go.  The training loop,
optimizer wrappers and dataset management are the reference's control plane and are not rebuilt.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple, Type, TypeVar, Union

import torch
from torch import Tensor

from . import __version__
from .attribute_graph import AttributeGraph, AttributeNode
from .config import Config, ProjectionEntryConfig
from .dataset_processing import Batch
from .loss_functions import LossWrapper
from .network.acoustic_model import Allophant, Predictions
from .phonetic_features import PhoneticAttributeIndexer
from .utils import evaluation

EstimatorCls = TypeVar("EstimatorCls", bound="Estimator")


def attribute_graph_from_config(config: Config, attribute_indexer: Any) -> AttributeGraph:
    """One node per configured classifier, sized by its number of categories (``run.py:212-226``)."""
    return AttributeGraph(
        AttributeNode(entry.name, attribute_indexer.size(entry.name), entry.time_layer, list(entry.dependencies))
        for entry in config.nn.projection.classes
    )


@dataclass
class Estimator:
    config: Config
    feature_size: int
    sample_rate: int
    attribute_graph: AttributeGraph
    model: Allophant
    loss_functions: Dict[str, LossWrapper]
    history: List[Any] = field(default_factory=list)
    epoch: Dict[str, int] = field(default_factory=lambda: {"epoch": 0, "global_step": 0, "step": 0})
    dataset_meta_data: List[Any] = field(default_factory=list)

    @classmethod
    def from_config(
        cls: Type[EstimatorCls],
        config: Config,
        feature_size: int,
        sample_rate: int,
        attribute_graph: AttributeGraph,
        attribute_indexer: Optional[PhoneticAttributeIndexer] = None,
        device: "torch.device | str" = "cuda",
        load_pretrained_weights: bool = True,
    ) -> EstimatorCls:
        model = Allophant.from_config(config.nn, feature_size, sample_rate, attribute_graph, attribute_indexer, load_pretrained_weights)
        model = model.to(device)
        return cls(config, feature_size, sample_rate, attribute_graph, model, config.nn.projection.loss_functions())

    @torch.inference_mode()
    def predict(
        self, batch: Batch, target_feature_indices: Optional[Tensor] = None, log_probabilities: bool = True, cuda_graph: bool = False
    ) -> Predictions:
        """``estimator.py:1035-1046``.  ``cuda_graph=True`` (an addition; log-probabilities only) replays the ~190 launches of the
        step as ONE captured CUDA graph per input shape: single utterances and small batches are bound by launch overhead
        (1 x 5 s: 2.9 ms eager, most of it host time), the graph removes it.  Results are copies, as fresh as eager ones."""
        with evaluation(self.model):
            if log_probabilities:
                if cuda_graph:
                    return self._predict_graphed(batch, target_feature_indices)
                return self.model.predict_log_probabilities(batch, target_feature_indices)
            return self.model(batch, target_feature_indices, predict=True)

    def _predict_graphed(self, batch: Batch, target_feature_indices: Optional[Tensor]) -> Predictions:
        from .engine import weight_generation

        audio = batch.audio_features
        if not audio.is_cuda:
            raise RuntimeError("allophant_b200 runs on CUDA only: move the batch to the GPU (`batch.to('cuda')`)")
        tfi = target_feature_indices
        parameters = getattr(self, "_graph_parameters", None)
        if parameters is None:
            parameters = self._graph_parameters = list(self.model.parameters())
        # the graph bakes in raw pointers to packed weights and workspaces: any weight change retires it
        version = (weight_generation(),) + tuple(p._version for p in parameters)
        # one graph per launch list: batches that pad up to the same length bucket (engine.bucket_samples) replay the same graph
        samples = audio.shape[1]
        acoustic = self.model.acoustic_model
        if getattr(acoustic, "bucket_frames", 0) > 1 and audio.dim() == 2 and hasattr(acoustic, "_model") and acoustic._model.config.feat_extract_norm == "layer":
            from .engine import bucket_samples

            samples, _ = bucket_samples(audio.shape[1], acoustic._model.config, acoustic.bucket_frames)
        key = ((audio.shape[0], samples) + tuple(audio.shape[2:]), audio.dtype, str(audio.device), None if tfi is None else (tfi.data_ptr(), tfi._version, tuple(tfi.shape)))
        graphs = self.__dict__.setdefault("_graphs", {})
        entry = graphs.get(key)
        if entry is not None and entry["version"] != version:
            entry = None
        if entry is None:
            padded = audio if samples == audio.shape[1] else torch.nn.functional.pad(audio, (0, samples - audio.shape[1]))
            static = Batch(padded.clone(), batch.lengths.clone(), batch.language_ids.clone())
            self.model.predict_log_probabilities(static, tfi)  # eager warm-up: plans, packed weights, composed embeddings, attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                captured = self.model.predict_log_probabilities(static, tfi)
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            entry = graphs[key] = dict(graph=graph, batch=static, predictions=captured, version=version)
        static = entry["batch"]
        static.audio_features[:, : audio.shape[1]].copy_(audio)
        if samples != audio.shape[1]:
            static.audio_features[:, audio.shape[1] :].zero_()
        static.lengths.copy_(batch.lengths)
        static.language_ids.copy_(batch.language_ids)
        entry["graph"].replay()
        captured = entry["predictions"]
        flat = captured._decode_cache["flat"]
        fresh = flat.clone()
        offset = flat.storage_offset()
        # the graph was captured on a batch of exactly the bucket's length; this batch may have fewer frames
        frames_shown = self.model.acoustic_model.downsampled_lengths(torch.tensor([audio.shape[1]])).item() if samples != audio.shape[1] else None
        outputs = {
            name: torch.as_strided(fresh, value.size(), value.stride(), value.storage_offset() - offset)[: frames_shown or value.shape[0]]
            for name, value in captured.outputs.items()
        }
        predictions = Predictions(outputs, captured.lengths.clone())
        cache = dict(captured._decode_cache)
        cache["flat"] = fresh
        for name in ("argmax", "maxlp", "frames32"):
            cache[name] = cache[name].clone()
        predictions._decode_cache = cache  # type: ignore[attr-defined]
        return predictions

    def map_allophones(self, phone_logits: Tensor, language_ids: Tensor) -> Tensor:
        return self.model.map_allophones(phone_logits, language_ids)

    # -- checkpoints --------------------------------------------------------------------------
    @staticmethod
    def _indexer_state(state: Any) -> Optional[Dict[str, Any]]:
        """``PhoneticIndexerState`` as the plain dictionary a checkpoint holds; an indexer is asked for its state."""
        if state is None or isinstance(state, dict):
            return state
        if hasattr(state, "state"):
            return state.state()
        allophones = state.language_allophones
        if allophones is not None and not isinstance(allophones, dict):
            allophones = {"allophones": allophones.allophones, "languages": allophones.languages, "shared_phones": allophones.shared_phones}
        return {"phoneme_inventory": list(state.phoneme_inventory), "language_allophones": allophones, "table_file": state.table_file}

    def checkpoint(
        self,
        phonetic_indexer_state: Any = None,
        optimizer_state: Optional[Dict[str, Any]] = None,
        additional_parameters: Optional[Dict[str, Any]] = None,
    ) -> Dict[str, Any]:
        """The checkpoint dictionary (``estimator.py:199-227``).  ``optimizer_state`` is the ``state_dict()`` of the optimizer
        (or of the ``OptimizerWrapper``, which includes the schedule's step); there is no gradient scaler (bf16 operands,
        fp32 master weights), so ``grad_scaler`` is always ``None``."""
        return {
            "config": self.config.dump(),
            "allophant_version": __version__,
            "feature_size": self.feature_size,
            "sample_rate": self.sample_rate,
            "attribute_graph": self.attribute_graph.state(),
            "epoch": dict(self.epoch),
            "phonetic_indexer_state": self._indexer_state(phonetic_indexer_state),
            "dataset_meta_data": list(self.dataset_meta_data),
            "model_state": self.model.state_dict(),
            "additional": additional_parameters,
            "history": list(self.history),
            "optimization_states": None if optimizer_state is None else {"optimizer": optimizer_state, "grad_scaler": None},
        }

    def save(
        self,
        file: Any,
        optimizer_state: Any = None,
        phonetic_indexer_state: Any = None,
        additional_parameters: Optional[Dict[str, Any]] = None,
    ) -> None:
        """``save(file, optimizer_state, phonetic_indexer_state, additional_parameters)`` as in ``estimator.py:1051-1083``.
        For convenience the indexer itself may be passed instead of its state, also in the second position
        (``save(file, indexer)``, the form earlier versions of this package used)."""
        if phonetic_indexer_state is None and optimizer_state is not None and hasattr(optimizer_state, "phonemes"):
            optimizer_state, phonetic_indexer_state = None, optimizer_state
        if phonetic_indexer_state is None:
            raise ValueError("a checkpoint needs the phonetic indexer state (the reference's Checkpoint schema requires it)")
        torch.save(self.checkpoint(phonetic_indexer_state, optimizer_state, additional_parameters), file)

    @classmethod
    def restore(
        cls: Type[EstimatorCls],
        checkpoint: Union[Dict[str, Any], Any],
        device: "torch.device | str" = "cuda",
        attribute_indexer: Optional[PhoneticAttributeIndexer] = None,
        **kwargs: Any,
    ) -> Tuple[EstimatorCls, PhoneticAttributeIndexer]:
        """Restores an estimator from a checkpoint dict, a local path or a file object.

        The reference also accepts a Hugging Face model id (``estimator.py:229-249``); there is no
        network here, so ids are treated as local paths."""
        if not isinstance(checkpoint, dict):
            checkpoint = torch.load(checkpoint, map_location=device, weights_only=True)
        config = Config.load(checkpoint["config"])
        composition_features = [
            entry.name for entry in config.nn.projection.classes if entry.name != ProjectionEntryConfig.PHONEME_LAYER
        ]
        if attribute_indexer is None:
            attribute_indexer = PhoneticAttributeIndexer.from_state(checkpoint["phonetic_indexer_state"], composition_features, config)
        graph = AttributeGraph.from_state(checkpoint["attribute_graph"])
        estimator = cls.from_config(
            config, checkpoint["feature_size"], checkpoint["sample_rate"], graph, attribute_indexer, device, load_pretrained_weights=False
        )
        estimator.model.load_state_dict(checkpoint["model_state"])
        estimator.epoch = dict(checkpoint.get("epoch") or estimator.epoch)
        estimator.history = list(checkpoint.get("history") or [])
        return estimator, attribute_indexer
