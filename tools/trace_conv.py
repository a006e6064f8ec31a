"""Debug aid: per-role clock64 timeline of one persistent CTA of the tcgen05 conv kernel (DMN_TC_TRACE=1).
slot = 16*tile_iter + k;  k: 0 producer tile start, 1 producer tables done, 2 producer last pass filled, 4 MMA got accumulators,
5 MMA first operand, 6 MMA tile issued, 8 epilogue tables done, 9 accumulators ready, 10 TMEM drained, 11 epilogue tile done"""
import ctypes as C
import os
import sys

os.environ["DMN_TC_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the production library compiles the timeline out (it costs 1.5 % of the step): use / build the trace variant
if "DMN_LIB_PATH" not in os.environ:
    _trace_lib = os.path.join(ROOT, "tools", "libdmn_trace.so")
    if not os.path.exists(_trace_lib):
        from diffusion_model_nemo_b200 import _build
        _build.build_variant(_trace_lib, ["-DDMN_TC_TRACE_BUILD=1", "-DDMN_EXP_MMATRACE=1"])
    os.environ["DMN_LIB_PATH"] = _trace_lib
sys.path.insert(0, os.path.join(ROOT, "tests"))
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
        kw = dict(gn=(8, torch.ones(cin).to(DEV), torch.zeros(cin).to(DEV)), silu=True)
        if os.environ.get("TRACE_TEMB"):      # per-image embedding rows (the layer API's form; the sampling loop shares one row)
            kw["temb"] = torch.zeros(b, cin).to(DEV)
    for _ in range(2):
        conv_forward(x.to(DEV), w.to(DEV), bias.to(DEV), mode=mode, ksize=k, out_groups=8 if mode == 0 and k == 3 else 0,
                     act=L.ACT_BF16, engine=L.CONV_TCGEN05, **kw)
    buf = (C.c_longlong * 1024)()
    lib.dmn_debug_conv_trace(buf, 1024)
    t = list(buf)
    if t[1000] and t[1001]:
        first = min(v for v in t[:960] if v > 10 ** 7)
        last = max(t[:960])
        print(f"   kernel entry -> first stamp {first - t[1000]} clk, last stamp -> exit {t[1001] - last} clk, entry -> exit {t[1001] - t[1000]} clk")
    t0 = min(v for v in t[:960] if v > 10 ** 7)
    rel = lambda v: (v - t0) if v else None
    per = [(t[16 * i + 6] - t[16 * i + 5]) for i in range(1, 10) if t[16 * i + 6] and t[16 * i + 16 + 6]]
    if per:
        print(f"   MMA main loop per tile (steady state): {sum(per) / len(per):.0f} clk")
    print(f"== k{k} mode{mode} {cin}->{cout} @{h}x{h} B={b} gn={gn}")
    for it in range(10):
        row = t[16 * it:16 * it + 16]
        if not any(row):
            break
        print(f"   tile {it}: prod start {rel(row[0])} tables {rel(row[1])} filled {rel(row[2])} | mma acc {rel(row[4])} firstA {rel(row[5])} "
              f"issued {rel(row[6])} (waitA {row[12]} waitB {row[13]}; prod waitEmpty {row[14]} waitCp {row[15]} issue {row[3]} finish {row[7]}) | epi tables {rel(row[8])} ready {rel(row[9])} drained {rel(row[10])} done {rel(row[11])}")
    if t[900]:
        print(f"   tile-2 setup (producer thread 0): enter {t[900]-t[32]} | first barrier {t[901]-t[900]} | pixel table {t[902]-t[901]} | gn table {t[903]-t[902]} | second barrier {t[16*2+1]-t[903]}")
    if t[910]:
        print(f"   tile-2 finish of pass 1 (producer thread 0): wait for the window {t[911]-t[910]} | coefficients {t[912]-t[911]} | items {t[913]-t[912]} | fence {t[914]-t[913]} | arrive {t[915]-t[914]}")
    pcs = t[800:800 + 96]
    if any(pcs):
        print("   epilogue pieces of tile 2 (warp 8): (wait, work) clk:", " ".join(
            f"({pcs[3*k+1]-pcs[3*k]},{pcs[3*k+2]-pcs[3*k+1]})" for k in range(32) if pcs[3 * k]))


if __name__ == "__main__" and os.environ.get("TRACE_SHAPES"):
    # TRACE_SHAPES="k,cin,cout,h,b,gn;..."
    for spec in os.environ["TRACE_SHAPES"].split(";"):
        k, cin, cout, h, b, gn, *rest = (int(v) for v in spec.split(","))
        run(k, cin, cout, h, b, bool(gn), mode=rest[0] if rest else 0)
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("TRACE_QUICK"):
    run(3, 128, 128, 32, 256, False)
    run(3, 256, 256, 16, 256, False)
    run(3, 128, 128, 32, 256, True)
    sys.exit(0)
if __name__ == "__main__":
    run(3, 128, 128, 32, 256, False)
    run(3, 128, 128, 32, 256, True)
    run(3, 256, 256, 16, 256, False)
    run(3, 256, 256, 16, 256, True)
    run(1, 128, 384, 32, 256, False)
    run(3, 256, 256, 4, 256, False)
    run(3, 256, 256, 4, 256, True)
    run(3, 256, 256, 8, 256, False)
    run(3, 256, 256, 8, 256, True)
