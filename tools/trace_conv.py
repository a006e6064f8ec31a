"""Debug aid: per-role clock64 timeline of one CTA of the tcgen05 conv kernel (DMN_TC_TRACE=1).

slots: 0 start, 1 setup done, 2 producers done, 3 accumulators ready, 4 epilogue done, 5 first B copy issued, 6 last B copy
issued, 7 CTA end; per pass c: 16+4c producer got the buffer, 17+4c producer filled it, 18+4c MMA got it, 19+4c MMA issued it."""
import ctypes as C
import os
import sys

os.environ["DMN_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from diffusion_model_nemo_b200 import _lib as L
from gpu_helpers import conv_forward

lib = L.lib()
lib.dmn_debug_conv_trace.argtypes = [C.c_void_p, C.c_int]
DEV = "cuda:0"


def run(k, cin, cout, h, b, gn, mode=0):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(b, cin, h, h, generator=g)
    w = torch.randn(*((cin, cout, 4, 4) if mode == 2 else (cout, cin, k, k)), generator=g) * 0.05
    bias = torch.zeros(cout)
    kw = {}
    if gn:
        kw = dict(gn=(8, torch.ones(cin).to(DEV), torch.zeros(cin).to(DEV)), silu=True, temb=torch.zeros(b, cin).to(DEV))
    for _ in range(2):
        conv_forward(x.to(DEV), w.to(DEV), bias.to(DEV), mode=mode, ksize=k, out_groups=8 if mode == 0 and k == 3 else 0,
                     act=L.ACT_BF16, engine=L.CONV_TCGEN05, **kw)
    buf = (C.c_longlong * 1024)()
    lib.dmn_debug_conv_trace(buf, 1024)
    t = list(buf)
    t0 = t[0]
    rel = lambda v: (v - t0) if v else None
    n_pass = (4 * cin if mode == 1 else cin) // 32
    print(f"== k{k} mode{mode} {cin}->{cout} @{h}x{h} B={b} gn={gn}: passes={n_pass}")
    print("   setup", rel(t[1]), " firstB", rel(t[5]), " lastB", rel(t[6]), " prod_done", rel(t[2]), " acc_ready", rel(t[3]),
          " epi_done", rel(t[4]), " end", rel(t[7]))
    print("   epilogue chunks (ld_issue, ld_done, stored, stats_done):")
    for i in range(8):
        print("     ", [rel(t[300 + i * 4 + k]) for k in range(4)])
    print("   per mt (chunks_done, scan_done, fence_done, bulk_issued):", [rel(t[400 + k]) for k in range(8)], " wait_group_done", rel(t[408]))
    for c in range(min(n_pass, 1)):
        print(f"   pass {c}: prod_get {rel(t[16+4*c])} prod_fill {rel(t[17+4*c])}  mma_get {rel(t[18+4*c])} mma_issued {rel(t[19+4*c])}")


if __name__ == "__main__":
    run(3, 128, 128, 32, 64, False)
    run(3, 256, 256, 4, 256, False)
    run(1, 128, 384, 32, 64, False)
