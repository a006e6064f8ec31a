"""The fused Residual(PreNorm(LinearAttention)) kernel (csrc/attn_fused.cu): TMA / SWIZZLE_128B plumbing self-test and block-level
parity against the oracle's restatement of reference utils.py:68-93 + parts/mha.py:44-59 (bf16 operands: rel-L2 <= 2e-2)."""
import pytest
import torch

from diffusion_model_nemo_b200 import _lib as L
from conftest import rel_l2
from gpu_helpers import linear_attention_block
from oracle import ref_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("M_,N,K", [(128, 128, 64), (256, 128, 128), (384, 256, 256), (200, 100, 192)])
def test_tma_sw128_gemm_selftest(M_, N, K):
    lib = L.lib()
    a, b = _bf(_rand(M_, K, seed=1)), _bf(_rand(N, K, seed=2))
    d = torch.zeros(M_, N, device=DEV)
    ad, bd = a.to(DEV, torch.bfloat16).contiguous(), b.to(DEV, torch.bfloat16).contiguous()
    L.check(lib.dmn_selftest_tma_sw128_gemm(L.ptr(ad), L.ptr(bd), L.ptr(d), M_, N, K, L.stream_ptr(DEV)), "selftest")
    assert rel_l2(d.cpu(), a @ b.T) <= 1e-5          # exact products, fp32 accumulation


def _block_sd(c, seed):
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, fan: (torch.rand(*s, generator=g) * 2 - 1) / fan ** 0.5      # noqa: E731
    p = "blk"
    return p, {
        p + ".fn.norm.weight": 1 + 0.1 * torch.randn(c, generator=g), p + ".fn.norm.bias": 0.1 * torch.randn(c, generator=g),
        p + ".fn.fn.to_qkv.weight": u(384, c, 1, 1, fan=c),
        p + ".fn.fn.to_out.0.weight": u(c, 128, 1, 1, fan=128), p + ".fn.fn.to_out.0.bias": 0.05 * torch.randn(c, generator=g),
        p + ".fn.fn.to_out.1.weight": 1 + 0.1 * torch.randn(c, generator=g), p + ".fn.fn.to_out.1.bias": 0.1 * torch.randn(c, generator=g),
    }


@pytest.mark.parametrize("b,c,h", [(1, 128, 16), (3, 128, 16), (2, 256, 16), (2, 128, 32), (5, 256, 32), (150, 128, 16),
                                   (3, 256, 8), (5, 256, 4), (2, 128, 8), (300, 256, 4), (1, 128, 4)])
def test_linear_attention_block_vs_oracle(b, c, h):
    p, sd = _block_sd(c, seed=7)
    x = _bf(_rand(b, c, h, h, seed=3) * 1.5 + 0.2)
    ref = O.residual_prenorm_fwd(sd, p, x, O.linear_attention_fwd)
    y = linear_attention_block(x.to(DEV), sd, p).cpu()
    assert torch.isfinite(y).all()
    # the residual x dominates the output; compare the attention branch itself as well
    assert rel_l2(y, ref) <= 1e-2
    assert rel_l2(y - x, ref - x) <= 2e-2


def test_linear_attention_block_large_k_logits():
    """Online column softmax: tiles whose maximum moves (sorted, growing logits) exercise the rescaling of the context accumulator."""
    b, c, h = 2, 128, 32
    p, sd = _block_sd(c, seed=9)
    sd[p + ".fn.fn.to_qkv.weight"] = sd[p + ".fn.fn.to_qkv.weight"] * 6.0
    ramp = torch.linspace(-2, 2, h * h).reshape(1, 1, h, h)
    x = _bf(_rand(b, c, h, h, seed=4) + ramp)
    ref = O.residual_prenorm_fwd(sd, p, x, O.linear_attention_fwd)
    y = linear_attention_block(x.to(DEV), sd, p).cpu()
    assert rel_l2(y - x, ref - x) <= 3e-2


@pytest.mark.parametrize("b,c,h", [(1, 256, 4), (5, 256, 4), (3, 128, 4), (2, 256, 8), (300, 256, 4)])
def test_softmax_attention_block_vs_oracle(b, c, h):
    """The bottleneck Residual(PreNorm(Attention)) (reference parts/mha.py:8-30) on the fused tcgen05 kernel."""
    p, sd = _block_sd(c, seed=11)
    sd[p + ".fn.fn.to_out.weight"] = sd.pop(p + ".fn.fn.to_out.0.weight")
    sd[p + ".fn.fn.to_out.bias"] = sd.pop(p + ".fn.fn.to_out.0.bias")
    sd[p + ".fn.fn.to_qkv.weight"] = sd[p + ".fn.fn.to_qkv.weight"] * 3.0      # logits of order 1: a non-trivial softmax
    x = _bf(_rand(b, c, h, h, seed=5) * 1.5 + 0.2)
    ref = O.residual_prenorm_fwd(sd, p, x, O.attention_fwd)
    y = linear_attention_block(x.to(DEV), sd, p, softmax=True).cpu()
    assert torch.isfinite(y).all()
    assert rel_l2(y, ref) <= 1e-2
    assert rel_l2(y - x, ref - x) <= 2e-2
