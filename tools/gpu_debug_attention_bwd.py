"""Debug helper: runs the attention backward sub-kernels one at a time (APH_ATT_BWD_MASK) with a sync after each."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mask = int(sys.argv[1]) if len(sys.argv) > 1 else 7
seq = int(sys.argv[2]) if len(sys.argv) > 2 else 249
os.environ["APH_ATT_BWD_MASK"] = str(mask)
import torch

from allophant_b200 import ops

DEV = "cuda"
torch.manual_seed(0)
lengths = [seq, max(1, seq // 2)]
n_utt, heads, d = len(lengths), 4, 64
hidden = heads * d
t_v = (seq + 7) // 8 * 8
scale = 0.125 * 1.4426950408889634
q = (torch.randn(n_utt, heads, seq, d, device=DEV) * scale).bfloat16()
k = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16()
v = torch.randn(n_utt, heads, seq, d, device=DEV).bfloat16()
vt = torch.zeros(n_utt, heads, d, t_v, device=DEV, dtype=torch.bfloat16)
vt[..., :seq] = v.transpose(-1, -2)
frames = torch.tensor(lengths, device=DEV, dtype=torch.int32)
ctx = torch.zeros(n_utt * seq, hidden, device=DEV, dtype=torch.bfloat16)
lse = torch.zeros(n_utt * heads * seq, device=DEV, dtype=torch.float32)
print("forward...", flush=True)
ops.attention(q, k, v, ctx, frames, n_utt, heads, seq, lse)
torch.cuda.synchronize()
print("forward ok", float(ctx.float().abs().mean()), float(lse.mean()), flush=True)
d_ctx = torch.randn(n_utt * seq, hidden, device=DEV).bfloat16()
dqkv = torch.zeros(n_utt * seq, 3 * hidden, device=DEV, dtype=torch.bfloat16)
delta = torch.empty(n_utt * heads * seq, device=DEV, dtype=torch.float32)
print(f"backward mask={mask}...", flush=True)
from allophant_b200 import _lib
progress = torch.zeros(16, dtype=torch.int32).pin_memory()
_lib.check(_lib.lib.aph_debug_set_progress(progress.data_ptr()), "dbg")
t0 = time.time()
ops.attention_backward(q, k, v, ctx, d_ctx, lse, delta, dqkv, frames, n_utt, heads, seq)
for _ in range(3):
    time.sleep(1.0)
    print("progress", progress.tolist(), flush=True)
torch.cuda.synchronize()
print(f"backward ok in {time.time() - t0:.3f}s", float(dqkv.float().abs().mean()), flush=True)
