"""Host-side cost of one streaming predict step (enqueue vs decode), and a cProfile of the enqueue path."""
import cProfile
import os
import pstats
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


def launch(timers):
    t0 = time.perf_counter()
    batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
    t1 = time.perf_counter()
    predictions = estimator.predict(batch, tfi_dev)
    t2 = time.perf_counter()
    pending = decode_predictions_async(predictions)
    t3 = time.perf_counter()
    timers["h2d"] += t1 - t0
    timers["predict"] += t2 - t1
    timers["decode_enqueue"] += t3 - t2
    return pending


for _ in range(3):
    launch({"h2d": 0, "predict": 0, "decode_enqueue": 0}).result()
torch.cuda.synchronize()
timers = {"h2d": 0.0, "predict": 0.0, "decode_enqueue": 0.0, "result_wait": 0.0, "result_host": 0.0}
steps = 10
start = time.perf_counter()
pending = None
for _ in range(steps):
    launched = launch(timers)
    if pending is not None:
        t0 = time.perf_counter()
        pending._event.synchronize()
        t1 = time.perf_counter()
        pending.result()
        t2 = time.perf_counter()
        timers["result_wait"] += t1 - t0
        timers["result_host"] += t2 - t1
    pending = launched
pending.result()
torch.cuda.synchronize()
total = time.perf_counter() - start
print(f"total {1000 * total / steps:.2f} ms/step")
for key, value in timers.items():
    print(f"  {key:16s} {1000 * value / steps:8.3f} ms/step")

profiler = cProfile.Profile()
profiler.enable()
for _ in range(5):
    launch({"h2d": 0, "predict": 0, "decode_enqueue": 0})
profiler.disable()
torch.cuda.synchronize()
pstats.Stats(profiler).sort_stats("cumulative").print_stats(35)
