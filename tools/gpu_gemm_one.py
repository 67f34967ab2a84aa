"""One FFN1-shaped GEMM (M=15968, N=4096, K=1024, bias+GELU -> bf16) for ncu captures."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from allophant_b200 import ops

M, N, K = 15968, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 1024
gelu = (len(sys.argv) > 3 and sys.argv[3] == "gelu")
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = torch.randn(N, device="cuda")
for _ in range(5):
    out = ops.linear_bf16(a, w, b, gelu=gelu)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.linear_bf16(a, w, b, gelu=gelu)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"M={M} N={N} K={K} gelu={gelu}: {ms:.4f} ms = {2*M*N*K/ms/1e9:.1f} TFLOP/s")
