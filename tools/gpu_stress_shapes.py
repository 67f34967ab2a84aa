"""GPU stress: odd batch shapes through Estimator.predict + greedy decode and one training step on the 2-layer test model —
tiny clips (one output frame), ragged lengths, single utterances, many short utterances.  Checks: no error, finite and
normalised log-probabilities on the valid frames, frame counts by the length formula, finite gradients."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allophant_b200 import predictions as decoding
from allophant_b200.config import Config, PhonemeLayerType
from allophant_b200.dataset_processing import Batch
from allophant_b200.estimator import Estimator, attribute_graph_from_config
from allophant_b200.loss_functions import multi_head_ctc_loss
from allophant_b200.network import wav2vec2
from allophant_b200.phonetic_features import PhoneticAttributeIndexer
import dataclasses

DEV = "cuda"


def frames_of(lengths):
    for kernel, stride in zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2)):
        lengths = torch.div(lengths - kernel, stride, rounding_mode="floor") + 1
    return lengths


def main():
    wav2vec2.KNOWN_MODELS["test/stress"] = dataclasses.replace(wav2vec2.KNOWN_MODELS["facebook/wav2vec2-xls-r-300m"], num_hidden_layers=2)
    config = Config.default()
    config.nn.projection.phoneme_layer = PhonemeLayerType.SHARED
    config.nn.acoustic_model.model_id = "test/stress"
    names = [entry.name for entry in config.nn.projection.classes]
    indexer = PhoneticAttributeIndexer.synthetic(80, names, training_inventory=50)
    graph = attribute_graph_from_config(config, indexer)
    torch.manual_seed(0)
    estimator = Estimator.from_config(config, 1, 16000, graph, indexer, DEV, load_pretrained_weights=False)
    tfi = indexer.composition_feature_matrix([f"p{i}" for i in range(30)]).to(DEV)
    cases = [[400], [401], [799, 400], [16000], [16000, 9000, 12345], [48000, 400, 800, 33333, 47999], [3200] * 40, [480000], [640, 1280, 1279]]
    for lengths in cases:
        lengths = torch.tensor(lengths)
        samples = int(lengths.max())
        audio = 0.1 * torch.randn(len(lengths), samples) * (torch.arange(samples)[None, :] < lengths[:, None])
        batch = Batch(audio.to(DEV), lengths.to(DEV), torch.zeros(len(lengths), dtype=torch.long, device=DEV))
        predictions = estimator.predict(batch, tfi)
        frames = frames_of(lengths)
        assert torch.equal(predictions.lengths.cpu(), frames), (lengths, predictions.lengths, frames)
        valid = (torch.arange(int(frames.max()))[:, None] < frames[None, :]).to(DEV)
        for name, value in predictions.outputs.items():
            assert bool(torch.isfinite(value[valid]).all()), (name, lengths)
            assert float(torch.logsumexp(value.float(), -1)[valid].abs().max()) < 1e-4, (name, lengths)
        hypotheses = decoding.decode_predictions(predictions)
        assert len(hypotheses["phoneme"]) == len(lengths)
        # one training step on the same shape (labels of one symbol where there is room)
        model = estimator.model
        # train() mode needs sequences of at least mask_time_length frames (SpecAugment raises like HF's _compute_mask_indices)
        model.train(int(frames.max()) >= 10)
        for parameter in model.parameters():
            parameter.grad = None
        outputs = model(batch)
        outputs.outputs.pop("phone", None)
        order = list(outputs.outputs)
        label_lengths = torch.clamp(frames // 4, min=0)
        labels = torch.ones(len(lengths), max(1, int(label_lengths.max())), dtype=torch.long)
        losses = multi_head_ctc_loss([outputs.outputs[n] for n in order], [labels.to(DEV)] * len(order), outputs.lengths, [label_lengths.to(DEV)] * len(order))
        (losses.sum() / max(1, int(label_lengths.sum()) * len(order))).backward()
        model.eval()
        bad = [name for name, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
        assert not bad, (lengths, bad[:3])
        torch.cuda.synchronize()
        print(f"ok {lengths.tolist()[:6]}{'...' if len(lengths) > 6 else ''} frames {frames.tolist()[:6]} loss {float(losses.sum()):.3f}", flush=True)
    print("stress ok")


if __name__ == "__main__":
    main()
