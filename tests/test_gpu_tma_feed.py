"""The two operand feeds of the tcgen05 conv engine must agree bit for bit: the TMA im2col feed (default for the hot 128-column
instantiations) and the cp.async producers it replaced (DMN_CONV_TMA=0) stage the same bf16 values, the tensor core accumulates the
same products in the same order, and the GroupNorm prologue applies the same fp32 arithmetic.  The switch is read once per process,
so every case runs in two child processes."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from diffusion_model_nemo_b200 import _lib as L
from gpu_helpers import conv_forward
DEV = "cuda:0"
def rnd(*s, seed=0): return torch.randn(*s, generator=torch.Generator().manual_seed(seed))
out = {}
# (name, mode, ksize, cin, cout, h, b, prologue): shapes that dispatch to the hot instantiations (N tile 128), incl. ragged last tiles
CASES = [("c3_l0", 0, 3, 128, 128, 32, 3, False), ("c3_gn", 0, 3, 128, 128, 32, 3, True), ("c3_256_gn", 0, 3, 256, 256, 16, 5, True),
         ("c3_small_gn", 0, 3, 256, 256, 4, 37, True), ("c3_8", 0, 3, 256, 128, 8, 9, False), ("down", 1, 4, 128, 128, 32, 3, False),
         ("down_small", 1, 4, 256, 256, 8, 5, False), ("up", 2, 4, 256, 256, 8, 5, False), ("up_l1", 2, 4, 128, 128, 16, 3, False),
         ("c1", 0, 1, 384, 128, 16, 3, False)]
for name, mode, k, cin, cout, h, b, pro in CASES:
    x = rnd(b, cin, h, h, seed=1)
    w = rnd(*((cin, cout, 4, 4) if mode == 2 else (cout, cin, k, k)), seed=2) / (cin * k * k) ** 0.5
    kw = {}
    if pro:
        kw = dict(gn=(8, (1 + 0.1 * rnd(cin, seed=4)).to(DEV), (0.1 * rnd(cin, seed=5)).to(DEV)), silu=True)
    y, st = conv_forward(x.to(DEV), w.to(DEV), (0.1 * rnd(cout, seed=3)).to(DEV), mode=mode, ksize=k,
                         out_groups=8 if (mode == 0 and k == 3) else 0, act=L.ACT_BF16, engine=L.CONV_TCGEN05, **kw)
    out[name] = y.cpu().numpy()
    if st is not None:
        out[name + "_stats"] = st.cpu().numpy()
np.savez(sys.argv[2], **out)
"""


def _run(tmp_path, tma, swap):
    path = str(tmp_path / f"feed_{tma}_{swap}.npz")
    env = dict(os.environ, DMN_CONV_TMA=tma, DMN_CONV_SWAP=swap)
    subprocess.run([sys.executable, "-c", CHILD, ROOT, path], check=True, env=env, timeout=600)
    return np.load(path)


def test_tma_feed_equals_cp_async_feed_bitwise(tmp_path):
    a, b = _run(tmp_path, "1", "0"), _run(tmp_path, "0", "0")
    assert sorted(a.files) == sorted(b.files) and len(a.files) >= 10
    for k in a.files:
        assert np.isfinite(a[k]).all(), k
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k] - b[k]).max()}"


def test_swapped_operand_roles_equal_plain_roles(tmp_path):
    """The swapped-role form (weights on the TMEM lanes, one M128 x N256 instruction per k-step) accumulates the same products in the
    same k order: the bf16 outputs are bit-identical; the GroupNorm statistics are summed in a different fp32 order (per channel over
    16-position blocks instead of per 32-row block), so they agree to rounding."""
    a, b = _run(tmp_path, "1", "1"), _run(tmp_path, "1", "0")
    for k in a.files:
        if k.endswith("_stats"):
            assert np.allclose(a[k], b[k], rtol=1e-5, atol=1e-6), k
        else:
            assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k] - b[k]).max()}"


UNET_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import make_unet
from oracle import ref_port as O
import diffusion_model_nemo_b200.modules as M
cfg = dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8)
u = make_unet(cfg, O.random_state_dict(cfg, seed=0), dtype="bf16", engine="tcgen05", device="cuda:0")
g = torch.Generator().manual_seed(5)
x = torch.randn(6, 3, 32, 32, generator=g).cuda()
t = torch.tensor([999, 500, 250, 3, 1, 0]).cuda()
eps = u(x, t)
# and a short free-running DDPM loop with injected noise (graph replay of the same kernels)
s = M.GaussianDiffusion(6, "linear")
noise = torch.stack([torch.randn(6, 3, 32, 32, generator=g) for _ in range(7)])
img = s.sample(u, [6, 3, 32, 32], device="cuda:0", noise=noise)[-1]
np.savez(sys.argv[2], eps=eps.float().cpu().numpy(), img=img.float().cpu().numpy())
"""


def test_whole_unet_and_loop_do_not_depend_on_the_operand_feed(tmp_path):
    """configs[1] U-Net (bf16, tcgen05) and a 6-step DDPM loop: TMA feed == cp.async feed bit for bit (same issue form in both)."""
    out = []
    for tma in ("1", "0"):
        path = str(tmp_path / f"unet_{tma}.npz")
        env = dict(os.environ, DMN_CONV_TMA=tma, DMN_CONV_SWAP="0")
        subprocess.run([sys.executable, "-c", UNET_CHILD, ROOT, path], check=True, env=env, timeout=900)
        out.append(np.load(path))
    for k in ("eps", "img"):
        assert np.isfinite(out[0][k]).all()
        assert np.array_equal(out[0][k], out[1][k]), f"{k}: max abs diff {np.abs(out[0][k] - out[1][k]).max()}"
