"""Supplementary measurement: the from-scratch pre-LN transformer acoustic model on one B200.

Workload: 80-dimensional features at 100 frames/s, 32 utterances x 10 s (1000 frames), linear frontend 80 -> 512, one GLU
convolution (kernel 3, stride 2: 500 frames into the transformer), 12 layers of width 512 / 8 heads / 2048 feed-forward
(GELU), Multitask heads (36 attribute classifiers + composed phoneme head), log_softmax + greedy decode of all 37 heads.
Prints one JSON line: audio-seconds per second (device-timed, inputs resident), kernel-only GEMM throughput from CUPTI."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from allophant_b200 import ops
from allophant_b200.config import Config, PhonemeLayerType, TransformerAcousticModelConfig
from allophant_b200.dataset_processing import Batch
from allophant_b200.estimator import Estimator, attribute_graph_from_config
from allophant_b200.phonetic_features import PhoneticAttributeIndexer

DEVICE = "cuda:0"
FEATURES, FRAMES_PER_SECOND, SECONDS, BATCH = 80, 100, 10, 32
LAYERS, WIDTH, HEADS, FF = 12, 512, 8, 2048


def main() -> None:
    torch.cuda.set_device(0)
    torch.manual_seed(2)
    config = Config.default()
    config.nn.projection.phoneme_layer = PhonemeLayerType.SHARED
    config.nn.acoustic_model = TransformerAcousticModelConfig.load(
        dict(
            type="pre-ln-transformer",
            transformer=dict(feedforward_neurons=FF, heads=HEADS, activation="gelu", num_layers=LAYERS, dropout_rate=0.1, positional_embeddings=True),
            frontend=dict(architecture="linear", neurons=WIDTH, input_dropout=0.0),
            sequential_frontend={"layers": [dict(type="glu1d", out_channels=WIDTH, kernel=3, stride=2), dict(type="layer_norm", affine=False)]},
            elementwise_affine=False,
        )
    )
    names = [entry.name for entry in config.nn.projection.classes]
    indexer = PhoneticAttributeIndexer.synthetic(109, names, n_categories=3, seed=1, training_inventory=60)
    graph = attribute_graph_from_config(config, indexer)
    estimator = Estimator.from_config(config, FEATURES, 16000, graph, indexer, device=DEVICE, load_pretrained_weights=False)
    tfi = indexer.composition_feature_matrix([f"p{index}" for index in range(25)]).to(DEVICE)
    frames = SECONDS * FRAMES_PER_SECOND
    features = torch.randn(BATCH, FEATURES, frames, device=DEVICE)
    lengths = torch.full((BATCH,), frames, dtype=torch.long, device=DEVICE)
    batch = Batch(features, lengths, torch.zeros(BATCH, dtype=torch.long, device=DEVICE))

    def step():
        predictions = estimator.predict(batch, tfi)
        cache = predictions._decode_cache
        return ops.ctc_greedy_collapse(cache["argmax"], cache["maxlp"], cache["frames32"], cache["n_utt"], cache["seq"], cache["argmax"].shape[0] * cache["n_utt"], 0)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ops.reset_launch_count()
    steps = 20
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        step()
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / steps
    launches = ops.launch_count() // steps
    rows = BATCH * (frames // 2)
    flops = LAYERS * (8 * WIDTH * WIDTH + 4 * WIDTH * FF) * rows
    measured = bench.cupti_gemm_time(step, lambda g: g.mode == 0 and g.n in (WIDTH, 3 * WIDTH, FF) and g.k in (WIDTH, FF))
    peaks = bench.measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    line = {
        "metric": "audio-sec/sec, pre-LN transformer acoustic model predict + greedy CTC decode (supplementary, SURVEY §8f rank 2)",
        "value": BATCH * SECONDS / (ms / 1000.0), "unit": "audio-s/s", "n_gpus": 1, "steps": steps, "ms_per_step": ms, "dtype": "bf16",
        "data": "synthetic", "gpu_launches": launches,
        "config": {"workload": f"{LAYERS} x (d {WIDTH}, {HEADS} heads, ff {FF}, GELU), linear frontend {FEATURES}->{WIDTH}, GLU conv k3 s2, {BATCH} x {SECONDS} s at {FRAMES_PER_SECOND} frames/s, 37 heads"},
    }  # fmt: skip
    if measured is not None:
        count, total_ms = measured
        achieved = flops / (total_ms / 1000.0) / 1e12
        line["roofline"] = {"kernel": "aph::gemm_bf16_kernel (transformer linears)", "bound": "tensor", "achieved": achieved, "peak": peak,
                            "frac": achieved / peak, "unit": "TFLOP/s", "launches": count, "avg_launch_ms": total_ms / count,
                            "share_of_step": total_ms / ms, "source": "CUPTI kernel records of one step"}  # fmt: skip
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
