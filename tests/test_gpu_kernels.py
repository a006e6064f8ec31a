"""GPU parity of the individual kernels, through the C ABI, against the CPU oracle (same seeded inputs)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from diffusion_model_nemo_b200 import _lib as L
import diffusion_model_nemo_b200.modules as M
from diffusion_model_nemo_b200.modules import sde as S
from conftest import rel_l2
from gpu_helpers import conv_forward, attention_core
from oracle import ref_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


# ---- sampler update kernels: <= 1e-6 relative (fp32, same eps in) ----------------------------------------------
@pytest.mark.parametrize("sched,T", [("linear", 1000), ("cosine", 250)])
def test_ddpm_step_kernel(sched, T):
    s = M.GaussianDiffusion(T, sched)
    tb = O.ddpm_tables(T, sched)
    x, eps, z = _rand(4, 3, 32, 32, seed=1), _rand(4, 3, 32, 32, seed=2), _rand(4, 3, 32, 32, seed=3)
    for ti in (T - 1, T // 2, 1, 0):
        t = torch.full((4,), ti, dtype=torch.long)
        ref = O.ddpm_step(tb, x, t, eps, z)
        got = s._fused_step(x.to(DEV), eps.to(DEV), torch.tensor([ti]), noise=z.to(DEV)).cpu()
        assert (got - ref).abs().max() <= 1e-6 * max(1.0, float(ref.abs().max())), ti
    # p_sample with a foreign (non-native) model callable
    model = lambda xx, tt: eps.to(DEV)            # noqa: E731
    got = s.p_sample(model, x.to(DEV), torch.full((4,), 7, device=DEV), noise=z.to(DEV)).cpu()
    assert (got - O.ddpm_step(tb, x, torch.full((4,), 7), eps, z)).abs().max() <= 1e-6 * 4


def test_learned_and_ddim_step_kernels():
    x, z = _rand(2, 3, 16, 16, seed=1), _rand(2, 3, 16, 16, seed=3)
    mo = _rand(2, 6, 16, 16, seed=2)
    s = M.LearnedGaussianDiffusion(250, "cosine")
    tb = O.ddpm_tables(250, "cosine")
    for ti in (249, 100, 0):
        ref = O.learned_step(tb, x, torch.full((2,), ti), mo, z)
        got = s._fused_step(x.to(DEV), mo.to(DEV), torch.tensor([ti]), noise=z.to(DEV)).cpu()
        assert (got - ref).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max()))
    eps = _rand(2, 3, 16, 16, seed=4)
    for eta in (0.0, 0.7):
        d = M.GeneralizedGaussianDiffusion(1000, "linear", eta=eta, ddim_timesteps=50)
        aext = O.ddim_extended_cumprod(O.ddpm_tables(1000, "linear")["betas"])
        for (ti, tn) in (d.timestep_pairs()[0], d.timestep_pairs()[20], d.timestep_pairs()[-1]):
            ref = O.ddim_step(aext, x, torch.full((2,), ti), torch.full((2,), tn), eps, z, eta)
            got, x0 = d.p_sample(lambda a, b: eps.to(DEV), x.to(DEV), torch.full((2,), ti, device=DEV),
                                 torch.full((2,), tn, device=DEV), noise=z.to(DEV))
            assert (got.cpu() - ref).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("kind", ["vp", "ve"])
def test_pc_update_kernels(kind):
    N = 1000
    sde = M.VPSDE(0.1, 20.0, N) if kind == "vp" else M.VESDE(0.01, 50.0, N)
    spec = O.SDESpec(kind, N=N)
    x, mo, z = _rand(3, 3, 32, 32, seed=1), _rand(3, 3, 32, 32, seed=2), _rand(3, 3, 32, 32, seed=3)
    cpu_model = lambda a, b: mo                   # noqa: E731
    gpu_model = lambda a, b: mo.to(DEV)           # noqa: E731
    sf = S.ScoreFunction(gpu_model, sde)
    ts = torch.linspace(1, spec.sampling_epsilon, N)
    for i in (0, 500, 999):
        vt = torch.ones(3) * ts[i]
        for name, fn, cls in (("reverse_diffusion", O.rd_predictor_step, M.ReverseDiffusionPredictor),
                              ("euler_maruyama", O.em_predictor_step, M.EulerMaruyamaPredictor)):
            xr, xm = fn(cpu_model, spec, x, vt, z)
            got, gm = cls(sde, sf).update_fn(x.to(DEV), vt.to(DEV), noise=z.to(DEV))
            tol = 4e-6 * max(1.0, float(xr.abs().max()))
            assert (got.cpu() - xr).abs().max() <= tol and (gm.cpu() - xm).abs().max() <= tol, (name, i)
        for name, fn, cls in (("langevin", O.langevin_step, M.LangevinCorrector), ("ald", O.ald_step, M.AnnealedLangevinDynamics)):
            xr, xm = fn(cpu_model, spec, x, vt, z, 0.16)
            got, gm = cls(sde, sf, 0.16, 1).update_fn(x.to(DEV), vt.to(DEV), noise=z.to(DEV)[None])
            tol = 1e-5 * max(1.0, float(xr.abs().max()))
            assert (got.cpu() - xr).abs().max() <= tol and (gm.cpu() - xm).abs().max() <= tol, (name, i)


def test_philox_normal_statistics_and_stream_independence():
    lib = L.lib()
    n = 1 << 22
    a = torch.empty(n, device=DEV)
    b = torch.empty(n, device=DEV)
    L.check(lib.dmn_randn(L.ptr(a), n, L.Rng(1234, 0), -1, L.stream_ptr(DEV)))
    L.check(lib.dmn_randn(L.ptr(b), n, L.Rng(1234, 1), -1, L.stream_ptr(DEV)))
    c = torch.empty(n, device=DEV)
    L.check(lib.dmn_randn(L.ptr(c), n, L.Rng(1234, 0), -1, L.stream_ptr(DEV)))
    assert torch.equal(a, c)                                       # deterministic in (seed, stream, step)
    assert abs(float(a.mean())) < 3e-3 and abs(float(a.var()) - 1) < 5e-3
    assert abs(float((a * b).mean())) < 3e-3                       # rank streams uncorrelated
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05                # kurtosis of a normal
    assert float(a.abs().max()) < 7.0 and torch.isfinite(a).all()


# ---- convolution engine (fp32 CUDA-core path): every mode and fusion against torch's own ops -----------------------
@pytest.mark.parametrize("mode,k,cin,cout,h", [(0, 3, 32, 64, 16), (0, 1, 64, 384, 8), (1, 4, 32, 32, 16), (2, 4, 64, 64, 8),
                                               (0, 3, 128, 128, 7), (2, 4, 32, 32, 7), (1, 4, 64, 64, 14)])
def test_conv_simt_fp32_vs_torch(mode, k, cin, cout, h):
    x = _rand(3, cin, h, h, seed=1)
    w = _rand(*((cin, cout, k, k) if mode == 2 else (cout, cin, k, k)), seed=2) / (cin * k * k) ** 0.5
    bias = _rand(cout, seed=3) * 0.1
    if mode == 0:
        ref = F.conv2d(x, w, bias, padding=k // 2)
    elif mode == 1:
        ref = F.conv2d(x, w, bias, stride=2, padding=1)
    else:
        ref = F.conv_transpose2d(x, w, bias, stride=2, padding=1)
    y, st = conv_forward(x.to(DEV), w.to(DEV), bias.to(DEV), mode=mode, ksize=k, out_groups=8)
    assert rel_l2(y.cpu(), ref) <= 2e-6
    g = ref.reshape(3, 8, -1)
    assert torch.allclose(st[..., 0].cpu(), g.mean(-1), atol=2e-5)
    assert torch.allclose(st[..., 1].cpu(), (g.var(-1, unbiased=False) + 1e-5).rsqrt(), rtol=2e-4)


@pytest.mark.parametrize("act,tol", [(L.ACT_F32, 3e-6), (L.ACT_BF16, 1.2e-2)])
def test_conv_prologue_gn_silu_temb(act, tol):
    """conv2 of a ResnetBlock: conv(SiLU(GN(h)) + temb)  (reference parts/convnext.py:35-41,81-83)."""
    b, c, h = 4, 64, 16
    x = _rand(b, c, h, h, seed=1) * 2 + 0.5
    w = _rand(64, c, 3, 3, seed=2) / (c * 9) ** 0.5
    bias, gamma, beta, temb = _rand(64, seed=3) * 0.1, 1 + 0.1 * _rand(c, seed=4), 0.1 * _rand(c, seed=5), _rand(b, c, seed=6)
    ref = F.conv2d(F.silu(F.group_norm(x, 8, gamma, beta)) + temb[:, :, None, None], w, bias, padding=1)
    y, _ = conv_forward(x.to(DEV), w.to(DEV), bias.to(DEV), gn=(8, gamma.to(DEV), beta.to(DEV)), silu=True, temb=temb.to(DEV), act=act)
    assert rel_l2(y.cpu(), ref) <= tol
    # PreNorm GroupNorm(1) without SiLU feeding a bias-free 1x1 (to_qkv)
    w1 = _rand(384, c, 1, 1, seed=7) / c ** 0.5
    ref = F.conv2d(F.group_norm(x, 1, gamma, beta), w1)
    y, _ = conv_forward(x.to(DEV), w1.to(DEV), None, ksize=1, gn=(1, gamma.to(DEV), beta.to(DEV)), act=act)
    assert rel_l2(y.cpu(), ref) <= tol


@pytest.mark.parametrize("act,tol", [(L.ACT_F32, 1e-5), (L.ACT_BF16, 1.5e-2)])
@pytest.mark.parametrize("n_side", [4, 7, 16, 32])
def test_attention_cores(act, tol, n_side):
    qkv = _rand(3, 384, n_side, n_side, seed=n_side)
    sd = {"a.to_qkv.weight": torch.eye(384).reshape(384, 384, 1, 1)}

    def lin_core(q):     # LinearAttention.forward between to_qkv and to_out (reference parts/mha.py:46-58)
        b, _, h, w = q.shape
        qq, kk, vv = [t.reshape(b, 4, 32, h * w) for t in q.chunk(3, dim=1)]
        qq = qq.softmax(dim=-2) * 32 ** -0.5
        kk = kk.softmax(dim=-1)
        ctx = torch.einsum("bhdn,bhen->bhde", kk, vv)
        return torch.einsum("bhde,bhdn->bhen", ctx, qq).reshape(b, 128, h, w)

    def sm_core(q):      # Attention.forward between to_qkv and to_out (reference parts/mha.py:18-29)
        b, _, h, w = q.shape
        qq, kk, vv = [t.reshape(b, 4, 32, h * w) for t in q.chunk(3, dim=1)]
        sim = torch.einsum("bhdi,bhdj->bhij", qq * 32 ** -0.5, kk)
        at = (sim - sim.amax(-1, keepdim=True)).softmax(-1)
        return torch.einsum("bhij,bhdj->bhid", at, vv).permute(0, 1, 3, 2).reshape(b, 128, h, w)

    assert rel_l2(attention_core(qkv.to(DEV), True, act).cpu(), lin_core(qkv)) <= tol
    if n_side <= 8:
        assert rel_l2(attention_core(qkv.to(DEV), False, act).cpu(), sm_core(qkv)) <= tol
