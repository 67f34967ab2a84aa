"""Host-side profile of one training step (cProfile) and GPU-busy vs wall time."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from allophant_b200.dataset_processing import Batch
from allophant_b200.loss_functions import multi_head_ctc_loss

device = "cuda:0"
torch.cuda.set_device(0)
estimator, allophones = bench.build_training_estimator(device)
model = estimator.model
model.train()
generator = torch.Generator().manual_seed(3)
seconds = 3.0 + 12.0 * torch.rand(bench.TRAIN_BATCH, generator=generator)
lengths = (seconds * bench.SAMPLE_RATE).long().sort(descending=True).values
samples = int(lengths.max())
audio = 0.1 * torch.randn(bench.TRAIN_BATCH, samples, generator=generator)
languages = torch.randint(0, bench.TRAIN_LANGUAGES, (bench.TRAIN_BATCH,), generator=generator)
frames = model.downsampled_lengths(lengths)
labels, label_lengths = {}, {}
for name in model.classes:
    classes = bench.TRAIN_PHONES + 1 if name == "phoneme" else 4
    head_lengths = (frames.double() * 0.25).floor().long()
    head_labels = torch.zeros(bench.TRAIN_BATCH, int(head_lengths.max()), dtype=torch.long)
    for row, length in enumerate(head_lengths.tolist()):
        if name == "phoneme":
            inventory = torch.tensor(sorted(allophones[int(languages[row])]), dtype=torch.long) + 1
            head_labels[row, :length] = inventory[torch.randint(0, len(inventory), (length,), generator=generator)]
        else:
            head_labels[row, :length] = torch.randint(1, classes, (length,), generator=generator)
    labels[name], label_lengths[name] = head_labels.to(device), head_lengths.to(device)
batch = Batch(audio.to(device), lengths.to(device), languages.to(device))
parameters = list(model.parameters())
from allophant_b200 import optim

optimizer = optim.adam_from_config(parameters, model.d_model, model=model)


def step(timers=None):
    t0 = time.perf_counter()
    for p in parameters:
        p.grad = None
    predictions = model(batch)
    predictions.outputs.pop("phone", None)
    order = list(predictions.outputs)
    t1 = time.perf_counter()
    losses = multi_head_ctc_loss([predictions.outputs[n] for n in order], [labels[n] for n in order], predictions.lengths, [label_lengths[n] for n in order])
    count = sum(label_lengths[n].sum() for n in order)
    loss = losses.sum() / count
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    optimizer.step(clip_norm=1.0)
    t4 = time.perf_counter()
    if timers is not None:
        timers["forward"] += t1 - t0
        timers["loss"] += t2 - t1
        timers["backward"] += t3 - t2
        timers["optimizer"] += t4 - t3
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
timers = {"forward": 0.0, "loss": 0.0, "backward": 0.0, "optimizer": 0.0}
start = time.perf_counter()
for _ in range(5):
    step(timers)
host_done = time.perf_counter()
torch.cuda.synchronize()
end = time.perf_counter()
print(f"host enqueue {1000 * (host_done - start) / 5:.2f} ms/step, wall {1000 * (end - start) / 5:.2f} ms/step")
for key, value in timers.items():
    print(f"  {key:10s} {1000 * value / 5:8.3f} ms/step (host)")
profiler = cProfile.Profile()
profiler.enable()
for _ in range(3):
    step()
profiler.disable()
torch.cuda.synchronize()
pstats.Stats(profiler).sort_stats("tottime").print_stats(30)

# ---- GPU timeline of one step through CUPTI (torch.profiler): busy time, gaps, top kernels
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
events = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
events.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in events)
span = events[-1].time_range.end - events[0].time_range.start
print(f"GPU events {len(events)}, busy {busy / 1000:.2f} ms, span {span / 1000:.2f} ms")
by_name = {}
for e in events:
    entry = by_name.setdefault(e.name[:70], [0, 0.0])
    entry[0] += 1
    entry[1] += e.time_range.end - e.time_range.start
for name, (count, total) in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{total / 1000:8.3f} ms {count:5d}x  {name}")
# largest idle gaps
gaps = []
for a, b in zip(events, events[1:]):
    gap = b.time_range.start - a.time_range.end
    if gap > 0:
        gaps.append((gap, a.name[:40], b.name[:40]))
gaps.sort(reverse=True)
print("total idle", sum(g[0] for g in gaps) / 1000, "ms; largest gaps:")
for gap, a, b in gaps[:12]:
    print(f"  {gap:8.1f} us between {a} -> {b}")

