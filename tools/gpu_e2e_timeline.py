"""GPU diagnostic: CUPTI timeline of bench.py's end-to-end predict stream (pinned host audio in, decoded tokens out).
Prints, for a few steady-state steps, what runs on which stream, the idle gaps of the compute stream and what surrounds them."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

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


def launch():
    with torch.cuda.stream(copy_stream):
        batch = Batch(host_audio, host_lengths, host_languages).to(device, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
    torch.cuda.current_stream().wait_event(copied)
    for tensor in (batch.audio_features, batch.lengths, batch.language_ids):
        tensor.record_stream(torch.cuda.current_stream())
    predictions = estimator.predict(batch, tfi_dev, cuda_graph=bench.E2E_CUDA_GRAPH)
    return decode_predictions_async(predictions)


def stream(steps):
    queue = []
    for _ in range(steps):
        queue.append(launch())
        if len(queue) > 2:
            queue.pop(0).result()
    while queue:
        queue.pop(0).result()


stream(6)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    stream(8)
    torch.cuda.synchronize()
events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
events.sort(key=lambda e: e.time_range.start)
first = events[0].time_range.start
starts = [e for e in events if "wave_stats" in e.name]
print(f"{len(events)} GPU activities, {len(starts)} steps; step starts (ms): " + ", ".join(f"{(e.time_range.start - first) / 1000:.2f}" for e in starts))
# one steady-state step: from the 4th wave_stats to the 5th
spacing = [(b.time_range.start - a.time_range.start) / 1000 for a, b in zip(starts, starts[1:])]
pick = min(range(len(spacing)), key=lambda i: spacing[i]) if spacing else 0  # the step the profiler's own host work disturbed least
if len(starts) >= 3:
    lo, hi = starts[pick].time_range.start, starts[pick + 1].time_range.start
    window = [e for e in events if lo <= e.time_range.start < hi]
    print(f"step window {(hi - lo) / 1000:.3f} ms, {len(window)} activities")
    non_kernel = [e for e in window if e.name.startswith("Mem")]
    for e in non_kernel:
        print(f"  {e.name[:40]:40s} at {(e.time_range.start - lo) / 1000:8.3f} ms, {(e.time_range.end - e.time_range.start):8.1f} us")
    kernels = [e for e in window if not e.name.startswith("Memcpy HtoD")]
    gaps = []
    for a, b in zip(kernels, kernels[1:]):
        gap = b.time_range.start - a.time_range.end
        if gap > 5:
            gaps.append((gap, (a.time_range.end - lo) / 1000, a.name[:44], b.name[:44]))
    print(f"  idle on the compute stream: {sum(g[0] for g in gaps) / 1000:.3f} ms in {len(gaps)} gaps > 5 us; largest:")
    for gap, at, a, b in sorted(gaps, reverse=True)[:14]:
        print(f"    {gap:7.1f} us at {at:7.3f} ms  {a} -> {b}")
    busy = sum(e.time_range.end - e.time_range.start for e in kernels)
    print(f"  busy {busy / 1000:.3f} ms")
