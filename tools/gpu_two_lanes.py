"""GPU experiment: does the end-to-end predict stream gain from TWO batches in flight on two CUDA streams (independent
workspaces), so that one batch's kernels fill the SMs the other's leave idle?  Two estimators stand in for two lanes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from allophant_b200.dataset_processing import Batch
from allophant_b200.predictions import decode_predictions_async

device = "cuda:0"
torch.cuda.set_device(0)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
estimators = []
for _ in range(lanes):
    estimator, tfi = bench.build_estimator(device)
    estimators.append((estimator, tfi.to(device)))
samples = bench.SECONDS * bench.SAMPLE_RATE
host_audio = (0.1 * torch.randn(bench.BATCH, samples)).pin_memory()
host_lengths = torch.full((bench.BATCH,), samples, dtype=torch.long).pin_memory()
host_languages = torch.zeros(bench.BATCH, dtype=torch.long).pin_memory()
copy_stream = torch.cuda.Stream(device=device)
streams = [torch.cuda.Stream(device=device) for _ in range(lanes)]


def launch(i):
    lane = i % lanes
    estimator, tfi_dev = estimators[lane]
    with torch.cuda.stream(copy_stream):
        batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
    with torch.cuda.stream(streams[lane]):
        streams[lane].wait_event(copied)
        for tensor in (batch.audio_features, batch.lengths, batch.language_ids):
            tensor.record_stream(streams[lane])
        predictions = estimator.predict(batch, tfi_dev, cuda_graph=True)
        return decode_predictions_async(predictions)


def stream(steps, depth):
    queue = []
    for i in range(steps):
        queue.append(launch(i))
        if len(queue) > depth:
            queue.pop(0).result()
    while queue:
        queue.pop(0).result()


for depth in (2, 3):
    stream(8, depth)
    torch.cuda.synchronize()
    steps = 80
    t0 = time.perf_counter()
    stream(steps, depth)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    print(f"lanes {lanes} depth {depth}: {wall * 1e3:.2f} ms/step, {bench.BATCH * bench.SECONDS / wall:.0f} audio-s/s")
