"""Debug helper: scaled vs exact CTC on bench-like shapes; prints flagged pairs and kernel times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allophant_b200 import ops

DEV = "cuda"
torch.manual_seed(0)
T, N = 612, 8
classes = [4] * 36 + [501]
if "narrow" in sys.argv[1:]:
    classes = [4] * 36
if "wide" in sys.argv[1:]:
    classes = [501]
input_lengths = torch.tensor([612, 600, 580, 500, 420, 300, 200, 150], device=DEV)
logits = [torch.randn(T, N, c, device=DEV) for c in classes]
if len(sys.argv) > 1 and sys.argv[1] == "masked":
    logits[-1][:, :, 250:] = torch.finfo(torch.float32).min  # absent phonemes of an allophone layer
log_probs = [ops.log_softmax(x) for x in logits]
labels, label_lengths = [], []
for c in classes:
    lengths = (input_lengths.double() * 0.25).floor().long()
    lab = torch.zeros(N, int(lengths.max()), dtype=torch.long, device=DEV)
    for n, length in enumerate(lengths.tolist()):
        lab[n, :length] = torch.randint(1, min(c, 250), (length,), device=DEV)
    labels.append(lab)
    label_lengths.append(lengths)
results = {}
for exact in (True,):
    problem = ops.CtcProblem(log_probs, labels, label_lengths, input_lengths, batch_first=False, need_grad=True)
    for _ in range(2):
        problem.forward()
        problem.backward(torch.ones(len(classes), device=DEV))
    torch.cuda.synchronize()
    start, mid, end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    start.record()
    problem.forward()
    mid.record()
    problem.backward(torch.ones(len(classes), device=DEV))
    end.record()
    torch.cuda.synchronize()
    print(f"exact={exact}: forward {start.elapsed_time(mid):.3f} ms, backward {mid.elapsed_time(end):.3f} ms")
    results[exact] = (problem.nll.clone(), [g.clone() for g in problem.grads])
# torch reference on the CPU for two heads
import torch.nn.functional as F
for head in sorted({0, len(classes) - 1}):
    lp = log_probs[head].detach().cpu().double().requires_grad_(True)
    loss = F.ctc_loss(lp, labels[head].cpu(), input_lengths.cpu(), label_lengths[head].cpu(), blank=0, reduction="none", zero_infinity=True)
    loss.sum().backward()
    ours = results[True][0][head].cpu().double()
    print(f"head {head}: nll rel err {float(((ours - loss.detach()).abs() / loss.detach().abs().clamp_min(1)).max()):.2e}")
