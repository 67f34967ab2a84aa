"""GPU bring-up check for the tcgen05 GEMM (run on the B200 box through gpurun).

Prints one line per case: max abs error against torch fp32 matmul on the same
bf16-rounded inputs, and a timing for the large shapes.
"""
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
from allophant_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def run(args):
    rc = L.lib.aph_gemm_bf16(ctypes.byref(args), stream())
    L.check(rc, "aph_gemm_bf16")
    torch.cuda.synchronize()


def plain(M, N, K, gelu=False, resid=False, bias=True, out="bf16", scale=1.0, lengths=None, period=0, label=""):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    r = torch.randn(M, N, device=dev) if resid else None
    ref = a.float() @ w.float().T * scale
    if bias:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r
    len_t = None
    if lengths is not None:
        len_t = torch.tensor(lengths, device=dev, dtype=torch.int32)
        rows = torch.arange(M, device=dev)
        mask = (rows % period) >= len_t[rows // period]
        ref[mask] = 0
    g = L.GemmArgs()
    g.a, g.a_row_stride, g.a_batch_stride, g.a_rows, g.a_inner, g.batch = ptr(a), K, 0, M, K, 1
    g.mode, g.b, g.n, g.k = L.APH_GEMM_ROWS, ptr(w), N, K
    g.epilogue, g.gelu, g.scale, g.bias = L.APH_EPI_STORE, int(gelu), scale, ptr(b)
    o32 = o16 = None
    if out in ("f32", "both"):
        o32 = torch.full((M, N), float("nan"), device=dev)
        if resid:
            o32.copy_(r)
            g.resid, g.ld_resid = ptr(o32), N  # in place
        g.out_f32, g.ld_f32 = ptr(o32), N
    elif resid:
        g.resid, g.ld_resid = ptr(r), N
    if out in ("bf16", "both"):
        o16 = torch.full((M, N), float("nan"), device=dev).bfloat16()
        g.out_bf16, g.ld_bf16 = ptr(o16), N
    g.out_batch_rows = 0
    if len_t is not None:
        g.lengths, g.len_period = ptr(len_t), period
    run(g)
    msgs = []
    for name, o in (("f32", o32), ("bf16", o16)):
        if o is None:
            continue
        err = (o.float() - ref).abs().max().item()
        rel = err / ref.abs().max().item()
        msgs.append(f"{name}: maxabs {err:.4g} rel {rel:.3g} nan {int(torch.isnan(o.float()).sum())}")
    print(f"[plain {label}] M={M} N={N} K={K} gelu={gelu} resid={resid} -> " + "; ".join(msgs), flush=True)
    return g, (a, w, b, r, o32, o16, len_t)


def conv_case(N_, L_in, C, kk, s):
    x = (torch.randn(N_, L_in, C, device=dev) * 0.5).bfloat16()
    w = (torch.randn(C, C, kk, device=dev) * 0.03).bfloat16()  # [out, in, k]
    b = torch.randn(C, device=dev)
    L_out = (L_in - kk) // s + 1
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), b, stride=s).transpose(1, 2)
    wp = w.permute(0, 2, 1).contiguous().view(C, kk * C)  # [o][j][c]
    out = torch.full((N_, L_out, C), float("nan"), device=dev).bfloat16()
    g = L.GemmArgs()
    g.a, g.a_row_stride, g.a_batch_stride = ptr(x), s * C, L_in * C
    g.a_rows, g.a_inner, g.batch = L_out, kk * C, N_
    g.mode, g.b, g.n, g.k = L.APH_GEMM_ROWS, ptr(wp), C, kk * C
    g.epilogue, g.gelu, g.scale, g.bias = L.APH_EPI_STORE, 0, 1.0, ptr(b)
    g.out_bf16, g.ld_bf16, g.out_batch_rows = ptr(out), C, L_out
    run(g)
    err = (out.float() - ref).abs().max().item()
    print(f"[conv] N={N_} L_in={L_in} k={kk} s={s} L_out={L_out}: maxabs {err:.4g} "
          f"rel {err / ref.abs().max().item():.3g} nan {int(torch.isnan(out.float()).sum())}", flush=True)


def taps_case(N_, T, C=1024, groups=16, taps=128):
    x = (torch.randn(N_, T, C, device=dev) * 0.5).bfloat16()
    cg = C // groups
    w = (torch.randn(C, cg, taps, device=dev) * 0.02).bfloat16()
    b = torch.randn(C, device=dev)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), b, padding=taps // 2, groups=groups)
    ref = ref[:, :, :-1].transpose(1, 2)
    ref = x.float() + torch.nn.functional.gelu(ref)
    wp = w.permute(0, 2, 1).contiguous().view(C, taps * cg)  # [o][tap][c]
    resid = x.float().contiguous()
    out = torch.full((N_, T, C), float("nan"), device=dev)
    g = L.GemmArgs()
    g.a, g.a_row_stride, g.a_batch_stride = ptr(x), C, T * C
    g.a_rows, g.a_inner, g.batch = T, C, N_
    g.mode, g.tap_pad, g.b, g.n, g.k = L.APH_GEMM_TAPS, taps // 2, ptr(wp), C, taps * cg
    g.epilogue, g.gelu, g.scale, g.bias = L.APH_EPI_STORE, 1, 1.0, ptr(b)
    g.resid, g.ld_resid = ptr(resid), C
    g.out_f32, g.ld_f32, g.out_batch_rows = ptr(out), C, T
    run(g)
    err = (out - ref).abs().max().item()
    print(f"[taps] N={N_} T={T}: maxabs {err:.4g} rel {err / ref.abs().max().item():.3g} "
          f"nan {int(torch.isnan(out).sum())}", flush=True)


def qkv_case(N_, T, heads=16):
    H = heads * 64
    M = N_ * T
    a = (torch.randn(M, H, device=dev) * 0.5).bfloat16()
    w = (torch.randn(3 * H, H, device=dev) * 0.03).bfloat16()
    b = torch.randn(3 * H, device=dev)
    ref = a.float() @ w.float().T + b
    tv = (T + 7) // 8 * 8
    q = torch.zeros(N_, heads, T, 64, device=dev).bfloat16()
    k = torch.zeros_like(q)
    vt = torch.zeros(N_, heads, 64, tv, device=dev).bfloat16()
    g = L.GemmArgs()
    g.a, g.a_row_stride, g.a_batch_stride, g.a_rows, g.a_inner, g.batch = ptr(a), H, 0, M, H, 1
    g.mode, g.b, g.n, g.k = L.APH_GEMM_ROWS, ptr(w), 3 * H, H
    g.epilogue, g.scale, g.bias = L.APH_EPI_QKV, 1.0, ptr(b)
    g.len_period, g.q, g.kmat, g.vt, g.heads, g.t_v, g.q_scale = T, ptr(q), ptr(k), ptr(vt), heads, tv, 0.125
    run(g)
    r = ref.view(N_, T, 3, heads, 64)
    eq = (q.float() - r[:, :, 0].permute(0, 2, 1, 3) * 0.125).abs().max().item()
    ek = (k.float() - r[:, :, 1].permute(0, 2, 1, 3)).abs().max().item()
    ev = (vt.float()[..., :T] - r[:, :, 2].permute(0, 2, 3, 1)).abs().max().item()
    print(f"[qkv] N={N_} T={T}: q {eq:.4g} k {ek:.4g} vt {ev:.4g} (ref max {ref.abs().max().item():.3g})", flush=True)


def timing(M, N, K, iters=20):
    g, keep = plain(M, N, K, gelu=False, resid=False, label="timing")
    for _ in range(3):
        L.lib.aph_gemm_bf16(ctypes.byref(g), stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.lib.aph_gemm_bf16(ctypes.byref(g), stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    a, w = keep[0], keep[1]
    for _ in range(3):
        torch.matmul(a, w.T)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, w.T)
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    print(f"[time] M={M} N={N} K={K}: ours {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s | "
          f"torch/cuBLAS {ms_t:.3f} ms = {fl / ms_t / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "abi", L.lib.aph_abi_version(), flush=True)
    t0 = time.time()
    plain(128, 256, 64, bias=False, label="1tile-1k")
    plain(128, 256, 1024, label="1tile")
    plain(1000, 1024, 1024, label="ragged-M")
    plain(4096, 1024, 1024, gelu=True, label="gelu")
    plain(2048, 1024, 4096, resid=True, out="both", label="resid-inplace")
    plain(1996, 784, 1024, out="f32", scale=0.5, label="heads-n784")
    plain(1996, 32, 640, out="f32", scale=1 / 640 ** 0.5, bias=False, label="compose-n32")
    plain(998, 96, 640, out="f32", bias=False, label="n96")
    plain(1996, 1024, 512, out="both", lengths=[499, 300, 1, 120], period=499, label="masked")
    conv_case(2, 1001, 512, 3, 2)
    conv_case(3, 3999, 512, 3, 2)
    conv_case(2, 999, 512, 2, 2)
    taps_case(2, 499)
    taps_case(1, 130)
    qkv_case(2, 499)
    timing(16384, 4096, 1024)
    timing(16384, 1024, 4096)
    timing(15968, 3072, 1024)
    print(f"done in {time.time() - t0:.1f}s", flush=True)
