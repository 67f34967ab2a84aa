"""GPU bring-up check for the tcgen05 varlen attention kernel."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from allophant_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def case(N, H, T, lengths, iters=0):
    tv = (T + 7) // 8 * 8
    q = (torch.randn(N, H, T, 64, device=dev)).bfloat16()
    k = (torch.randn(N, H, T, 64, device=dev)).bfloat16()
    v = (torch.randn(N, H, T, 64, device=dev)).bfloat16()
    vt = torch.zeros(N, H, 64, tv, device=dev).bfloat16()
    vt[..., :T] = v.transpose(2, 3)
    qs = (q.float() * 0.125 * 1.4426950408889634).bfloat16()
    ctx = torch.zeros(N * T, H * 64, device=dev).bfloat16()
    len_t = torch.tensor(lengths, device=dev, dtype=torch.int32)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.lib.aph_attention_bf16(ptr(qs), ptr(k), ptr(vt), ptr(ctx), ptr(len_t), N, H, T, tv, st)
    L.check(rc, "aph_attention_bf16")
    torch.cuda.synchronize()
    mask = torch.arange(T, device=dev)[None, :] < len_t[:, None]  # [N, T]
    s = (qs.float() / 1.4426950408889634 @ k.float().transpose(2, 3))
    s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    ref = torch.softmax(s, -1) @ v.float()  # [N,H,T,64]
    ref = ref.permute(0, 2, 1, 3).reshape(N, T, H * 64)
    out = ctx.float().view(N, T, H * 64)
    err = 0.0
    for n in range(N):
        ln = lengths[n]
        if ln:
            err = max(err, (out[n, :ln] - ref[n, :ln]).abs().max().item())
    print(f"[att] N={N} H={H} T={T} len={lengths[:6]}: maxabs {err:.4g} (ref max {ref.abs().max().item():.3g}) "
          f"nan {int(torch.isnan(out).sum())}", flush=True)
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            L.lib.aph_attention_bf16(ptr(qs), ptr(k), ptr(vt), ptr(ctx), ptr(len_t), N, H, T, tv, st)
        e0.record()
        for _ in range(iters):
            L.lib.aph_attention_bf16(ptr(qs), ptr(k), ptr(vt), ptr(ctx), ptr(len_t), N, H, T, tv, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = sum(4.0 * H * 64 * l * l for l in lengths)
        qq, kk, vv = q.clone(), k, v
        for _ in range(3):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e0.record()
        for _ in range(iters):
            torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / iters
        print(f"[att time] {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s | torch sdpa (no mask) {ms_t:.3f} ms", flush=True)


if __name__ == "__main__":
    case(1, 1, 128, [128])
    case(1, 2, 100, [100])
    case(2, 16, 499, [499, 300])
    case(3, 4, 749, [749, 1, 130])
    case(2, 2, 1499, [1499, 1000])
    case(32, 16, 499, [499] * 32, iters=20)
    case(8, 16, 1499, [1499] * 8, iters=10)
