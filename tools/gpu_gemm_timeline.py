import os, sys
sys.path.insert(0, "/root/repo")
import torch
from allophant_b200 import ops
DEV = "cuda"
torch.manual_seed(0)
m, h, ff = 15968, 1024, 4096
x = (torch.randn(m, h, device=DEV) * 0.5).bfloat16()
w1 = (torch.randn(ff, h, device=DEV) * 0.03).bfloat16()
b1 = torch.randn(ff, device=DEV)
act = torch.zeros(m, ff, device=DEV, dtype=torch.bfloat16)
for gelu in (True, False):
    args = ops.make_gemm_args(x, w1, a_rows=m, a_inner=h, a_row_stride=h, bias=b1, gelu=gelu, out_bf16=act, ld_bf16=ff)
    for _ in range(3):
        ops.run_gemm(args)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.run_gemm(args)
    e.record(); torch.cuda.synchronize()
    print(f"gelu={gelu}: {s.elapsed_time(e) * 100:.1f} us per launch")
    tl = torch.zeros(512, device=DEV, dtype=torch.int64)
    os.environ["APH_GEMM_TIMELINE"] = str(tl.data_ptr())
    ops.run_gemm(args)
    torch.cuda.synchronize()
    del os.environ["APH_GEMM_TIMELINE"]
    t = tl.tolist()
    base = min(v for v in t if v)
    for tile in range(3):
        o = tile * 40
        print(f" tile {tile}: epi start +{t[o]-base}, tfull +{t[o+1]-base}, end +{t[o+38]-base}")
        for c in range(4):
            q = o + 2 + c * 8
            print(f"   chunk {c}: ld issue +{t[q]-base}  ld done {t[q+1]-t[q]}  bias {t[q+2]-t[q+1]}  gelu {t[q+3]-t[q+2]}  misc {t[q+4]-t[q+3]}")
        mo = 300 + tile * 40
        stamps = [t[mo + 2 * k] - base for k in range(16)]
        print(f"   mma: tempty +{t[mo+39]-base}; k-block waits done at", stamps)
