"""CUPTI timeline of one predict step: GPU busy time vs span, per-kernel totals, largest gaps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from allophant_b200 import ops
from allophant_b200.dataset_processing import Batch

device = "cuda:0"
torch.cuda.set_device(0)
estimator, tfi = bench.build_estimator(device)
tfi_dev = tfi.to(device)
samples = bench.SECONDS * bench.SAMPLE_RATE
batch = Batch((0.1 * torch.randn(bench.BATCH, samples)).to(device), torch.full((bench.BATCH,), samples, dtype=torch.long).to(device),
              torch.zeros(bench.BATCH, dtype=torch.long).to(device))


def step():
    predictions = estimator.predict(batch, tfi_dev)
    cache = predictions._decode_cache
    return ops.ctc_greedy_collapse(cache["argmax"], cache["maxlp"], cache["frames32"], cache["n_utt"], cache["seq"], cache["argmax"].shape[0] * cache["n_utt"], 0)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
events = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in events)
span = events[-1].time_range.end - events[0].time_range.start
print(f"3 steps: GPU events {len(events)}, busy {busy / 3000:.3f} ms/step, span {span / 3000:.3f} ms/step, idle {(span - busy) / 3000:.3f} ms/step")
by_name = {}
for e in events:
    entry = by_name.setdefault(e.name[:60], [0, 0.0])
    entry[0] += 1
    entry[1] += e.time_range.end - e.time_range.start
for name, (count, total) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"{total / 3000:8.3f} ms/step {count // 3:4d}x  avg {total / count:7.1f} us  {name}")
gaps = sorted(((b.time_range.start - a.time_range.end, a.name[:30], b.name[:30]) for a, b in zip(events, events[1:])), reverse=True)
print("largest gaps (us):", [(round(g[0], 1), g[1], g[2]) for g in gaps[:6]])
import statistics
print("median gap us:", statistics.median(g[0] for g in gaps))
