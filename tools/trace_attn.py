"""Phase timeline of the fused LinearAttention kernel (CTA 0): DMN_FA_TRACE=1 python tools/trace_attn.py [C] [H] [B]"""
import ctypes as C
import os
import sys

os.environ["DMN_FA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from diffusion_model_nemo_b200 import _lib as L  # noqa: E402
from gpu_helpers import linear_attention_block  # noqa: E402

c, h, b = (int(sys.argv[1]) if len(sys.argv) > 1 else 128), (int(sys.argv[2]) if len(sys.argv) > 2 else 32), (int(sys.argv[3]) if len(sys.argv) > 3 else 256)
g = torch.Generator().manual_seed(0)
p = "blk"
sd = {p + ".fn.norm.weight": torch.ones(c), p + ".fn.norm.bias": torch.zeros(c), p + ".fn.fn.to_qkv.weight": torch.randn(384, c, 1, 1, generator=g) / c ** 0.5,
      p + ".fn.fn.to_out.0.weight": torch.randn(c, 128, 1, 1, generator=g) / 128 ** 0.5, p + ".fn.fn.to_out.0.bias": torch.zeros(c),
      p + ".fn.fn.to_out.1.weight": torch.ones(c), p + ".fn.fn.to_out.1.bias": torch.zeros(c)}
x = torch.randn(b, c, h, h, generator=g).cuda()
for _ in range(2):
    linear_attention_block(x, sd, p)
lib = L.lib()
lib.dmn_debug_fa_trace.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_longlong * 16)()
lib.dmn_debug_fa_trace(buf, 16)
names = ["start", "phaseA", "ctx", "phaseB", "phaseC"]
for base, tag in ((0, "first image"), (8, "last image")):
    t = [buf[base + i] for i in range(5)]
    if t[0] == 0:
        continue
    print(tag, " ".join(f"{names[i]}=+{t[i] - t[i - 1]}" for i in range(1, 5)), f"total={t[4] - t[0]} clk")
