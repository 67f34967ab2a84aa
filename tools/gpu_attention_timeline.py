"""Debug helper: clock64 timeline of CTA (0,0) of the attention forward kernel on the bench shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allophant_b200 import _lib, ops

DEV = "cuda"
n_utt, heads, seq, d = 32, 16, 499, 64
t_v = (seq + 7) // 8 * 8
q = (torch.randn(n_utt * heads, seq, d, device=DEV) * 0.125 * 1.4427).bfloat16()  # bench-like: no rescaling after block 0
k = torch.randn(n_utt * heads, seq, d, device=DEV).bfloat16()
v = torch.randn(n_utt * heads, seq, d, device=DEV).bfloat16()
ctx = torch.zeros(n_utt * seq, heads * d, device=DEV, dtype=torch.bfloat16)
frames = torch.full((n_utt,), seq, device=DEV, dtype=torch.int32)
for _ in range(3):
    ops.attention(q, k, v, ctx, frames, n_utt, heads, seq)
torch.cuda.synchronize()
start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for _ in range(20):
    ops.attention(q, k, v, ctx, frames, n_utt, heads, seq)
end.record()
torch.cuda.synchronize()
print(f"attention kernel: {start.elapsed_time(end) / 20 * 1000:.1f} us per launch")
timeline = torch.zeros(192, device=DEV, dtype=torch.int64)
_lib.check(_lib.lib.aph_debug_set_timeline(timeline.data_ptr()), "timeline")
EXPERIMENT = int(os.environ.get("APH_ATT_EXPERIMENT", "0"))  # timeline builds: bit 0 PV cut, 1 scores cut, 2 no exponentials
if EXPERIMENT:
    import ctypes

    _lib.check(
        _lib.lib.aph_attention_bf16_dropout(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), None, frames.data_ptr(), n_utt, heads, seq, 0, EXPERIMENT, ctypes.c_float(1.0), None
        ),
        "attention",
    )
else:
    ops.attention(q, k, v, ctx, frames, n_utt, heads, seq)
torch.cuda.synchronize()
stamps = timeline.tolist()
base = stamps[0]
names = {0: "entry", 1: "alloc+sync done", 2: "Q loaded (mma)", 24: "last PV done", 25: "epilogue done", 26: "dealloc done"}
for j in range(10):
    names[4 + 2 * j] = f"S_{j} ready"
    names[5 + 2 * j] = f"P_{j} arrived"
names.update({27: "  blk2: scores loaded", 28: "  blk2: max + any done", 29: "  blk2: exp2 + sums done", 30: "  blk2: P buffer free", 31: "  blk2: P stored"})
for slot in sorted(names, key=lambda s: stamps[s]):
    if stamps[slot]:
        print(f"{names[slot]:18s} +{stamps[slot] - base:7d} clk")

# steady-state detail of the pair kernel's third item on CTA 0 (only in a library built with EXTRA=-DAPH_ATT_TIMELINE)
if any(stamps[32:]):
    detail = {}
    for t, tile in enumerate("AB"):
        for j in range(4):
            base = 32 + t * 32 + j * 8
            detail[base + 0] = f"softmax {tile}: S({j}) seen"
            detail[base + 1] = f"softmax {tile}: S({j}) released"
            detail[base + 2] = f"softmax {tile}: keys 64-127 of ({j}) stored"
            detail[base + 5] = f"softmax {tile}: keys 0-63 of ({j}) loaded"
            detail[base + 6] = f"softmax {tile}: maximum of ({j}) known"
            detail[base + 7] = f"softmax {tile}: P columns of ({j}) free"
            detail[base + 3] = f"softmax {tile}: P({j}) store issued"
            detail[base + 4] = f"softmax {tile}: P({j}) arrived"
            detail[96 + j * 8 + t * 4 + 0] = f"mma {tile}: S({j + 1}) issued (or none)"
            detail[96 + j * 8 + t * 4 + 1] = f"mma {tile}: P({j}) seen"
            detail[96 + j * 8 + t * 4 + 2] = f"mma {tile}: PV({j}) issued"
        detail[160 + t * 4 + 0] = f"softmax {tile}: last PV seen"
        detail[160 + t * 4 + 1] = f"softmax {tile}: O read, TMEM handed back"
        detail[160 + t * 4 + 2] = f"softmax {tile}: context rows written"
    live = [slot for slot in detail if stamps[slot]]
    first = min(stamps[slot] for slot in live)
    for slot in sorted(live, key=lambda s: stamps[s]):
        print(f"{detail[slot]:40s} +{stamps[slot] - first:7d} clk")

# experiments of a timeline build: parts of the work switched off through the (otherwise unused) dropout seed
if any(stamps[32:]):
    import ctypes

    _lib.check(_lib.lib.aph_debug_set_timeline(0), "timeline off")
    for code, what in ((0, "full"), (1, "PV products: 1 of 8 steps"), (2, "score products: 1 of 4 steps"), (3, "both products cut"), (4, "no exponentials"), (7, "no exponentials, products cut")):
        def run():
            _lib.check(
                _lib.lib.aph_attention_bf16_dropout(
                    q.data_ptr(), k.data_ptr(), v.data_ptr(), ctx.data_ptr(), None, frames.data_ptr(), n_utt, heads, seq, 0, code, ctypes.c_float(1.0), None
                ),
                "attention",
            )
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        start.record()
        for _ in range(20):
            run()
        end.record()
        torch.cuda.synchronize()
        print(f"experiment [{what}]: {start.elapsed_time(end) / 20 * 1000:.1f} us per launch")
