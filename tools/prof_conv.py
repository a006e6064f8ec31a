"""One conv launch of the tcgen05 engine at bench shape, for `ncu --set full -k regex:conv_tcgen05`.
usage: python tools/prof_conv.py [plain|gn|small|qkv]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from diffusion_model_nemo_b200 import _lib as L
from gpu_helpers import conv_forward

DEV = "cuda:0"
CASES = {"plain": (3, 128, 128, 32, 256, False), "gn0": (3, 128, 128, 32, 256, True), "gn": (3, 256, 256, 16, 256, True), "small": (3, 256, 256, 4, 256, True),
         "qkv": (1, 128, 384, 32, 256, False)}


def main(which):
    k, cin, cout, h, b, gn = CASES[which]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(b, cin, h, h, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * 0.05
    kw = {}
    if gn:
        kw = dict(gn=(8, torch.ones(cin).to(DEV), torch.zeros(cin).to(DEV)), silu=True)
        if os.environ.get("PROF_TEMB"):       # per-image embedding rows (the layer API's form; the sampling loop shares one row)
            kw["temb"] = torch.zeros(b, cin).to(DEV)
    for _ in range(3):
        conv_forward(x.to(DEV), w.to(DEV), torch.zeros(cout).to(DEV), ksize=k, out_groups=8 if k == 3 else 0, act=L.ACT_BF16,
                     engine=L.CONV_TCGEN05, **kw)
    torch.cuda.synchronize()
    print("ok", which)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "plain")
