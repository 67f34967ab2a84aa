"""Epilogue-cost experiment: the encoder's four GEMM shapes with different epilogues (CUDA events, 30 launches each,
operands rotated through 4 buffer sets so that inputs do not sit in L2 more than in the real step)."""
import sys

import torch

sys.path.insert(0, ".")
from allophant_b200 import ops

M = 15968
DEV = "cuda"


def bench(name, n, k, *, out="bf16", resid=False, gelu=False, sets=4, reps=30):
    a = [(torch.randn(M, k, device=DEV) * 0.5).bfloat16() for _ in range(sets)]
    w = [(torch.randn(n, k, device=DEV) * 0.05).bfloat16() for _ in range(sets)]
    b = torch.randn(n, device=DEV)
    o32 = [torch.randn(M, n, device=DEV) for _ in range(sets)] if out == "f32" else None
    o16 = [torch.empty(M, n, device=DEV, dtype=torch.bfloat16) for _ in range(sets)] if out == "bf16" else None
    args = []
    for i in range(sets):
        args.append(
            ops.make_gemm_args(
                a[i], w[i], a_rows=M, a_inner=k, a_row_stride=k, bias=b, gelu=gelu,
                resid=o32[i] if resid else None, ld_resid=n,
                out_f32=o32[i] if out == "f32" else None, ld_f32=n, out_bf16=o16[i] if out == "bf16" else None, ld_bf16=n,
            )
        )
    for i in range(sets):
        ops.run_gemm(args[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        ops.run_gemm(args[r % sets])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} N={n:5d} K={k:5d} out={out:4s} resid={int(resid)} gelu={int(gelu)}: {ms * 1000:7.1f} us  {2 * M * n * k / ms / 1e9:7.1f} TFLOP/s")


bench("out-proj (as in the encoder)", 1024, 1024, out="f32", resid=True)
bench("out-proj, fp32 out, no residual", 1024, 1024, out="f32")
bench("out-proj, bf16 out", 1024, 1024, out="bf16")
bench("FFN2 (as in the encoder)", 1024, 4096, out="f32", resid=True)
bench("FFN2, bf16 out", 1024, 4096, out="bf16")
bench("FFN1 (as in the encoder)", 4096, 1024, out="bf16", gelu=True)
bench("FFN1, no GELU", 4096, 1024, out="bf16")
bench("QKV-shaped plain bf16", 3072, 1024, out="bf16")
