"""Times the four encoder GEMM forms of BASELINE configs[1] (32 x 10 s: 15 968 rows) in one process: QKV scatter, out-proj
(+bias +fp32 residual in place), FFN1 (+bias +GELU), FFN2 (+bias +residual), and the LayerNorm-folded forms of each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allophant_b200 import ops

DEV = "cuda"
torch.manual_seed(0)
m, h, ff, heads, seq = 15968, 1024, 4096, 16, 499
x = (torch.randn(m, h, device=DEV) * 0.5).bfloat16()
act = (torch.randn(m, ff, device=DEV) * 0.5).bfloat16()
wqkv, bqkv = (torch.randn(3 * h, h, device=DEV) * 0.03).bfloat16(), torch.randn(3 * h, device=DEV)
wo, bo = (torch.randn(h, h, device=DEV) * 0.03).bfloat16(), torch.randn(h, device=DEV)
w1, b1 = (torch.randn(ff, h, device=DEV) * 0.03).bfloat16(), torch.randn(ff, device=DEV)
w2, b2 = (torch.randn(h, ff, device=DEV) * 0.03).bfloat16(), torch.randn(h, device=DEV)
q, k, v = (torch.zeros(32 * heads * seq * 64, device=DEV, dtype=torch.bfloat16) for _ in range(3))
hidden = torch.randn(m, h, device=DEV)
copy16 = torch.zeros(m, h, device=DEV, dtype=torch.bfloat16)
out_ff = torch.zeros(m, ff, device=DEV, dtype=torch.bfloat16)
stats = torch.rand(m, 8, 2, device=DEV)
colsum3, colsum4 = torch.randn(3 * h, device=DEV), torch.randn(ff, device=DEV)
flush = torch.empty(256 * 1024 * 1024, device=DEV, dtype=torch.uint8)


def timed(name, args, flops):
    for _ in range(3):
        ops.run_gemm(args)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.run_gemm(args)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1000)
    times.sort()
    us = times[len(times) // 2]
    print(f"{name:34s} {us:7.1f} us  {flops / us / 1e6:7.1f} TFLOP/s")


plain = dict(a_rows=m, a_inner=h, a_row_stride=h)
timed("QKV", ops.make_qkv_args(x, wqkv, bqkv, q, k, v, rows=m, seq=seq, heads=heads), 2 * m * h * 3 * h)
timed("QKV + folded LN", ops.with_layernorm(ops.make_qkv_args(x, wqkv, bqkv, q, k, v, rows=m, seq=seq, heads=heads), stats, colsum3, h, 1e-5), 2 * m * h * 3 * h)
timed("out-proj (+res)", ops.make_gemm_args(x, wo, bias=bo, resid=hidden, ld_resid=h, out_f32=hidden, ld_f32=h, **plain), 2 * m * h * h)
timed("out-proj (+res +stats +copy)", ops.with_row_stats(ops.make_gemm_args(x, wo, bias=bo, resid=hidden, ld_resid=h, out_f32=hidden, ld_f32=h, out_bf16=copy16, ld_bf16=h, **plain), stats), 2 * m * h * h)
timed("FFN1 (+GELU)", ops.make_gemm_args(x, w1, bias=b1, gelu=True, out_bf16=out_ff, ld_bf16=ff, **plain), 2 * m * h * ff)
timed("FFN1 (+GELU) + folded LN", ops.with_layernorm(ops.make_gemm_args(x, w1, bias=b1, gelu=True, out_bf16=out_ff, ld_bf16=ff, **plain), stats, colsum4, h, 1e-5), 2 * m * h * ff)
timed("FFN1 without GELU", ops.make_gemm_args(x, w1, bias=b1, out_bf16=out_ff, ld_bf16=ff, **plain), 2 * m * h * ff)
timed("FFN2 (+res)", ops.make_gemm_args(act, w2, a_rows=m, a_inner=ff, a_row_stride=ff, bias=b2, resid=hidden, ld_resid=h, out_f32=hidden, ld_f32=h), 2 * m * h * ff)
timed("FFN2 (+res +stats +copy)", ops.with_row_stats(ops.make_gemm_args(act, w2, a_rows=m, a_inner=ff, a_row_stride=ff, bias=b2, resid=hidden, ld_resid=h, out_f32=hidden, ld_f32=h, out_bf16=copy16, ld_bf16=h), stats), 2 * m * h * ff)
