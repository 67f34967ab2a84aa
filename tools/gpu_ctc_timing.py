"""GPU diagnostic: time of the CTC alpha / beta launches against the number of heads, the class count and the frame count
(one warp per (utterance, head) pair walks time sequentially: the launch takes as long as its slowest warp)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allophant_b200 import ops

dev = "cuda"


def case(n_heads, n_utt, frames, classes, label_fraction, ragged):
    torch.manual_seed(0)
    lengths = torch.full((n_utt,), frames, dtype=torch.long)
    if ragged:
        lengths = torch.linspace(frames // 4, frames, n_utt).long()
    label_len = int(frames * label_fraction)
    log_probs, labels, label_lengths = [], [], []
    for _ in range(n_heads):
        log_probs.append(torch.randn(frames, n_utt, classes, device=dev).log_softmax(-1))
        labels.append(torch.randint(1, classes, (n_utt, label_len), device=dev))
        label_lengths.append((lengths.double() * label_fraction).floor().long().to(dev))
    problem = ops.CtcProblem(log_probs, labels, label_lengths, lengths.to(dev), batch_first=False, need_grad=True)
    scale = torch.ones(n_heads, device=dev)
    for _ in range(2):
        problem.forward()
        problem.backward(scale)
    torch.cuda.synchronize()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 5
    fwd = bwd = 0.0
    for _ in range(reps):
        events[0].record()
        problem.forward()
        events[1].record()
        problem.backward(scale)
        events[2].record()
        torch.cuda.synchronize()
        fwd += events[0].elapsed_time(events[1])
        bwd += events[1].elapsed_time(events[2])
    fwd, bwd = 1e3 * fwd / reps, 1e3 * bwd / reps
    print(f"heads {n_heads:3d} utt {n_utt:3d} frames {frames:4d} classes {classes:4d} labels {label_len:4d} ragged {int(ragged)}: "
          f"alpha {fwd:7.1f} us ({fwd * 1.9e3 / frames:6.0f} cyc/frame)  beta {bwd:7.1f} us ({bwd * 1.9e3 / frames:6.0f} cyc/frame)")


CASES = [
    (1, 4, 613, 4, 0.25, False),
    (1, 8, 613, 4, 0.25, False),
    (8, 8, 613, 4, 0.25, False),
    (37, 8, 613, 4, 0.25, False),
    (37, 8, 613, 4, 0.25, True),
    (37, 8, 613, 4, 0.05, False),
    (1, 8, 613, 501, 0.25, False),
    (1, 8, 613, 26, 0.25, False),
    (74, 8, 613, 4, 0.25, False),
    (37, 64, 613, 4, 0.25, False),
]
for index, args in enumerate(CASES):
    if len(sys.argv) < 2 or int(sys.argv[1]) == index:
        case(*args)
