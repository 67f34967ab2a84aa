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
timeline = torch.zeros(32, device=DEV, dtype=torch.int64)
_lib.check(_lib.lib.aph_debug_set_timeline(timeline.data_ptr()), "timeline")
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
