"""GPU diagnostic: where the end-to-end predict stream loses time against the device-timed step.
Runs bench.py's streaming client at pipeline depths 1-3 and reports wall ms/step, plus the host time spent in launch /
wait / hypothesis assembly per step."""
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
estimator, tfi = bench.build_estimator(device)
tfi_dev = tfi.to(device)
samples = bench.SECONDS * bench.SAMPLE_RATE
host_audio = (0.1 * torch.randn(bench.BATCH, samples)).pin_memory()
host_lengths = torch.full((bench.BATCH,), samples, dtype=torch.long).pin_memory()
host_languages = torch.zeros(bench.BATCH, dtype=torch.long).pin_memory()
copy_stream = torch.cuda.Stream(device=device)
timers = {"launch": 0.0, "wait+host": 0.0}


def launch():
    t0 = time.perf_counter()
    with torch.cuda.stream(copy_stream):
        batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
    torch.cuda.current_stream().wait_event(copied)
    for tensor in (batch.audio_features, batch.lengths, batch.language_ids):
        tensor.record_stream(torch.cuda.current_stream())
    predictions = estimator.predict(batch, tfi_dev)
    pending = decode_predictions_async(predictions)
    timers["launch"] += time.perf_counter() - t0
    return pending


def stream(steps, depth):
    queue = []
    for _ in range(steps):
        queue.append(launch())
        if len(queue) > depth:
            t0 = time.perf_counter()
            queue.pop(0).result()
            timers["wait+host"] += time.perf_counter() - t0
    while queue:
        queue.pop(0).result()


resident = Batch(host_audio, host_lengths, host_languages).to(device)
for _ in range(3):
    estimator.predict(resident, tfi_dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    estimator.predict(resident, tfi_dev)
torch.cuda.synchronize()
print(f"predict only (resident inputs, no decode): {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/step")
for depth in (1, 2):
    stream(4, depth)
    torch.cuda.synchronize()
    timers["launch"] = timers["wait+host"] = 0.0
    steps = 60
    t0 = time.perf_counter()
    stream(steps, depth)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    print(f"depth {depth}: {wall:.2f} ms/step wall; host launch {timers['launch'] / steps * 1e3:.2f} ms, wait+assembly {timers['wait+host'] / steps * 1e3:.2f} ms")
# the same stream without the host hypothesis assembly: only the device work + copies
from allophant_b200 import predictions as P

for depth in (1,):
    steps = 60
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    queue = []
    for _ in range(steps):
        queue.append(launch())
        if len(queue) > depth:
            queue.pop(0)._event.synchronize()
    torch.cuda.synchronize()
    print(f"depth {depth}, no host assembly: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms/step wall")
