"""Times the memory-bound kernels at BASELINE config sizes and reports achieved GB/s (algorithmic bytes)."""
import sys
import torch

sys.path.insert(0, ".")
from allophant_b200 import ops
from allophant_b200.loss_functions import multi_head_ctc_loss

dev = "cuda"


def synthetic_labels(frames, n_classes, seed, fraction=0.25):
    """Random label sequences in [1, n_classes) of length floor(fraction * frames)."""
    generator = torch.Generator().manual_seed(seed)
    lengths = (frames.double() * fraction).floor().long().clamp_min(1)
    labels = torch.zeros(len(frames), int(lengths.max()), dtype=torch.long)
    for row, length in enumerate(lengths.tolist()):
        labels[row, :length] = torch.randint(1, n_classes, (length,), generator=generator)
    return labels, lengths

PEAK = 6542.7


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms*1000:9.1f} us  {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured HBM peak", flush=True)


# config 4: wide phoneme head, 128 x 499 frames x 3184 classes
rows, width = 128 * 499, 3184
logits = torch.randn(rows, width, device=dev)
out = torch.empty_like(logits)
am = torch.empty(rows, dtype=torch.int32, device=dev)
mx = torch.empty(rows, device=dev)
report("log_softmax_wide [63872 x 3184] (+argmax)", timeit(lambda: ops.log_softmax_wide(logits, width, rows, width, out, width, am, mx)), 2 * rows * width * 4)
report("  torch.log_softmax same shape", timeit(lambda: torch.log_softmax(logits, -1)), 2 * rows * width * 4)

# config 2: 36 narrow heads packed [15968 x 144]
rows2 = 32 * 499
packed = torch.randn(rows2, 144, device=dev)
col = torch.arange(0, 144, 4, dtype=torch.int32, device=dev)
wid = torch.full((36,), 4, dtype=torch.int32, device=dev)
off = (torch.arange(36, dtype=torch.int64, device=dev) * rows2 * 4)
out2 = torch.empty(rows2 * 144, device=dev)
am2 = torch.empty(36, rows2, dtype=torch.int32, device=dev)
mx2 = torch.empty(36, rows2, device=dev)
report("log_softmax_heads 36 heads [15968 x 144] (+argmax)", timeit(lambda: ops.log_softmax_heads(packed, 144, rows2, 0, 144, col, wid, off, 36, out2, am2, mx2)), 2 * rows2 * 144 * 4 + 36 * rows2 * 8)

# layernorm fp32 [15968 x 1024] -> bf16
x = torch.randn(rows2, 1024, device=dev)
g, b = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
o16 = torch.empty(rows2, 1024, device=dev, dtype=torch.bfloat16)
report("layernorm_rows fp32->bf16 [15968 x 1024]", timeit(lambda: ops.layernorm_rows(x, rows2, 1024, 1024, g, b, 1e-5, out_bf16=o16, ld_bf16=1024)), rows2 * 1024 * 6)
# layernorm+gelu bf16 in place [32*15999 x 512]
rows3 = 32 * 15999
xb = torch.randn(rows3, 512, device=dev).bfloat16()
g5, b5 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
report("layernorm_rows+gelu bf16 in place [511968 x 512]", timeit(lambda: ops.layernorm_rows(xb, rows3, 512, 512, g5, b5, 1e-5, gelu=True, out_bf16=xb, ld_bf16=512), iters=10), rows3 * 512 * 4)
del xb

# conv0 fused: 32 x 160000 samples -> [32, 31999, 512] bf16
audio = torch.randn(32, 160000, device=dev) * 0.1
lengths = torch.full((32,), 160000, dtype=torch.int64, device=dev)
stats = torch.empty(32, 3, dtype=torch.float64, device=dev)
mr = torch.empty(32, 2, device=dev)
ops.wave_stats(audio, lengths, stats, mr)
w0 = torch.randn(512, 10, device=dev) * 0.3
o0 = torch.empty(32, 31999, 512, device=dev, dtype=torch.bfloat16)
report("conv0+norm+LN+GELU [32 x 160000] -> [32,31999,512] bf16", timeit(lambda: ops.conv0_ln_gelu(audio, lengths, mr, w0, g5, g5, b5, 1e-5, o0), iters=10), 32 * 160000 * 4 + 32 * 31999 * 512 * 2)
report("wave_stats [32 x 160000]", timeit(lambda: ops.wave_stats(audio, lengths, stats, mr)), 32 * 160000 * 4)
del o0

# CTC: config 3 shape per GPU (8 utterances) and the whole batch (64), 36 heads c=4 + phoneme c=501, T'=749
for n_utt in (8, 64):
    frames = 749
    input_lengths = torch.randint(150, frames + 1, (n_utt,))
    input_lengths[0] = frames
    classes = [4] * 36 + [501]
    logits_l, labels_l, lens_l = [], [], []
    for h, c in enumerate(classes):
        logits_l.append(torch.randn(n_utt, frames, c, device=dev).transpose(0, 1).requires_grad_(True))
        lab, ll = synthetic_labels(input_lengths, c, seed=h)
        labels_l.append(lab.to(dev))
        lens_l.append(ll.to(dev))
    il = input_lengths.to(dev)
    log_probs = [ops.log_softmax(t.detach()) for t in logits_l]
    problem = ops.CtcProblem(log_probs, labels_l, lens_l, il, batch_first=False, need_grad=True)
    scale = torch.ones(len(classes), device=dev)
    valid = int(input_lengths.sum())
    lp_bytes = sum(valid * c * 4 for c in classes)
    alpha_bytes = valid * problem.s_pad * 4 * len(classes)
    report(f"ctc_forward  N={n_utt} T'=749 37 heads (alpha stored, S_pad={problem.s_pad})", timeit(problem.forward, iters=10), lp_bytes + alpha_bytes)
    report(f"ctc_backward N={n_utt} (beta + grad)", timeit(lambda: problem.backward(scale), iters=10), 2 * lp_bytes + alpha_bytes)
    def torch_ctc():
        total = 0
        for lp, lab, ll in zip(log_probs, labels_l, lens_l):
            total = total + torch.nn.functional.ctc_loss(lp, lab, il, ll, reduction="sum", zero_infinity=True)
        return total
    report(f"  torch ctc_loss forward x37 (no .item()) N={n_utt}", timeit(torch_ctc, iters=5), lp_bytes)

# greedy collapse: 37 heads x 32 utt x 499
am3 = torch.randint(0, 4, (37, rows2), dtype=torch.int32, device=dev)
mx3 = torch.randn(37, rows2, device=dev)
fl = torch.full((32,), 499, dtype=torch.int32, device=dev)
report("ctc_greedy_collapse 37 x 32 x 499", timeit(lambda: ops.ctc_greedy_collapse(am3, mx3, fl, 32, 499, 37 * 32, 0)), 37 * rows2 * 8)
