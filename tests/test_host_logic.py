"""Host side of the product (no GPU): bit-exact schedule tables, timestep indexing, per-step coefficient rows,
parameter naming, error behaviour.  The oracle is the checker here."""
import functools

import numpy as np
import pytest
import torch

import diffusion_model_nemo_b200.modules as M
from diffusion_model_nemo_b200 import _lib as L
from diffusion_model_nemo_b200.modules import sde as S
from diffusion_model_nemo_b200.modules import _runtime as R
from conftest import CFGS, make_unet
from oracle import ref_port as O


@pytest.mark.parametrize("name", ["linear", "quadratic", "sigmoid", "cosine"])
@pytest.mark.parametrize("T", [50, 250, 1000])
def test_tables_bit_exact_vs_oracle_and_golden(golden, name, T):
    s = M.GaussianDiffusion(timesteps=T, schedule_name=name)
    ref = O.ddpm_tables(T, name)
    for k in O.DDPM_TABLE_NAMES:
        t = getattr(s, k)
        assert t.dtype == torch.float32 and t.device.type == "cpu"
        assert torch.equal(t, ref[k]), (name, T, k)
    assert len(s.state_dict()) == 0        # tables are plain attributes, as in the reference


def test_schedule_cfg_and_errors(golden):
    s = M.GaussianDiffusion(100, "linear", schedule_cfg={"linear": {"beta_start": 1e-3, "beta_end": 0.05}})
    assert torch.equal(s.betas, O.ddpm_tables(100, "linear", {"linear": {"beta_start": 1e-3, "beta_end": 0.05}})["betas"])
    with pytest.raises(AssertionError):
        M.GaussianDiffusion(10, "exponential")
    with pytest.raises(AssertionError):
        M.GaussianDiffusion(10, "linear", objective="pred_v")
    with pytest.raises(ValueError):
        M.GeneralizedGaussianDiffusion(10, "linear", eta=1.5)
    with pytest.raises(ValueError):
        M.GaussianDiffusion(10, "linear").interpolate(None, torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4), t=10)
    with pytest.raises(RuntimeWarning):
        M.GaussianDiffusion(10, "linear")(1)


def test_ddim_tables_and_pairs(golden):
    for name, T, Sx in (("cosine", 1000, 50), ("linear", 1000, 10), ("linear", 20, 5)):
        s = M.GeneralizedGaussianDiffusion(T, name, eta=0.0, ddim_timesteps=Sx)
        assert torch.equal(s.alphas_extended_cumprod, O.ddim_extended_cumprod(O.ddpm_tables(T, name)["betas"]))
        assert s.timestep_pairs() == O.ddim_pairs(T, Sx)
    assert M.GeneralizedGaussianDiffusion(1000, "linear").ddim_timesteps == 1000   # ddim_timesteps=-1 -> T


def test_sde_tables_and_time_grid():
    vp, ve = M.VPSDE(0.1, 20.0, 1000), M.VESDE(0.01, 50.0, 1000)
    for k, v in O.vp_tables(0.1, 20.0, 1000).items():
        assert torch.equal(getattr(vp, k), v)
    assert torch.equal(ve.discrete_sigmas, O.ve_tables(0.01, 50.0, 1000)["discrete_sigmas"])
    s = M.PredictorCorrectorSampler("reverse_diffusion", "langevin", snr=0.16)
    with pytest.raises(ValueError):
        s.forward(None, [1, 3, 8, 8], "cuda")
    s.update_sde(vp)
    ts = s.timesteps()
    assert torch.equal(ts, torch.linspace(1, 1e-3, 1000))
    idx = (ts * 999 / 1).long()
    assert idx[0] == 999 and idx[-1] == 0 and bool((idx[:-1] >= idx[1:]).all())
    s.update_sde(ve)
    assert float(s.timesteps()[-1]) == pytest.approx(1e-5)
    assert M.get_predictor("reverse_diffusion") is M.ReverseDiffusionPredictor
    assert M.get_corrector("langevin") is M.LangevinCorrector and M.get_corrector("nope") is None
    with pytest.raises(ValueError):
        M.register_predictor(M.ReverseDiffusionPredictor, "reverse_diffusion")


def _emulate(rows, s):
    return [float(c[s]) for c in rows]


@pytest.mark.parametrize("sched", ["linear", "cosine"])
def test_ddpm_rows_reproduce_oracle_step(sched):
    """The coefficient rows + the kernel's formula (restated in numpy) == oracle.ddpm_step."""
    T = 50
    s = M.GaussianDiffusion(T, sched)
    tb = O.ddpm_tables(T, sched)
    g = torch.Generator().manual_seed(0)
    x, eps, z = (torch.randn(2, 3, 8, 8, generator=g) for _ in range(3))
    ts = s._visit_order()
    assert ts.tolist() == list(reversed(range(T)))
    rows = s._step_rows(ts)
    for step in (0, 17, T - 2, T - 1):
        c0, c1, c2, c3, c4, flag = _emulate(rows, step)
        t = torch.full((2,), int(ts[step]), dtype=torch.long)
        x0 = (np.float32(c0) * x - np.float32(c1) * eps).clamp(-1, 1)
        mine = np.float32(c2) * x0 + np.float32(c3) * x + np.float32(c4) * z
        ref = O.ddpm_step(tb, x, t, eps, z)
        assert (mine - ref).abs().max() <= 1e-6 * max(1.0, float(ref.abs().max()))
    assert float(rows[4][-1]) == 0.0        # no noise at t = 0


def test_ddim_and_learned_rows():
    T = 20
    g = torch.Generator().manual_seed(0)
    x, eps, z = (torch.randn(2, 3, 8, 8, generator=g) for _ in range(3))
    for eta in (0.0, 0.5, 1.0):
        s = M.GeneralizedGaussianDiffusion(T, "linear", eta=eta, ddim_timesteps=5)
        pairs = s.timestep_pairs()
        t = torch.tensor([p[0] for p in pairs])
        tn = torch.tensor([p[1] for p in pairs])
        rows = s._pair_rows(t, tn)
        aext = O.ddim_extended_cumprod(O.ddpm_tables(T, "linear")["betas"])
        for k in range(len(pairs)):
            s1m, sa, san, k1, k2, _ = _emulate(rows, k)
            x0 = ((x - eps * np.float32(s1m)) / np.float32(sa)).clamp(-1, 1)
            mine = np.float32(san) * x0 + np.float32(k1) * z + np.float32(k2) * eps
            ref = O.ddim_step(aext, x, torch.full((2,), pairs[k][0]), torch.full((2,), pairs[k][1]), eps, z, eta)
            assert (mine - ref).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max()))
    s = M.LearnedGaussianDiffusion(T, "cosine")
    rows = s._step_rows(s._visit_order())
    tb = O.ddpm_tables(T, "cosine")
    mo = torch.randn(2, 6, 8, 8, generator=g)
    for step in (0, 7, T - 1):
        c0, c1, c2, c3, mask, lo, hi = _emulate(rows, step)
        e, v = mo.chunk(2, dim=1)
        frac = (v + 1) * 0.5
        lv = frac * np.float32(hi) + (1 - frac) * np.float32(lo)
        x0 = (np.float32(c0) * x - np.float32(c1) * e).clamp(-1, 1)
        mine = np.float32(c2) * x0 + np.float32(c3) * x + np.float32(mask) * torch.exp(0.5 * lv) * z
        ref = O.learned_step(tb, x, torch.full((2,), T - 1 - step), mo, z)
        assert (mine - ref).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("kind", ["vp", "ve"])
@pytest.mark.parametrize("pred", ["reverse_diffusion", "euler_maruyama"])
def test_pc_rows_reproduce_oracle_updates(kind, pred):
    N = 40
    sde = M.VPSDE(0.1, 20.0, N) if kind == "vp" else M.VESDE(0.01, 50.0, N)
    spec = O.SDESpec(kind, N=N)
    g = torch.Generator().manual_seed(1)
    x, mo, z = (torch.randn(3, 3, 8, 8, generator=g) for _ in range(3))
    model = lambda x_, t_: mo                     # noqa: E731  (a fixed "network output")
    sf = S.ScoreFunction(model, sde)
    ts = torch.linspace(1, spec.sampling_epsilon, N)
    a, b, gg = S.predictor_rows(pred, sde, sf, ts)
    kind_c, (sc, al) = S.corrector_rows("langevin", sde, sf, ts, 0.16)
    assert kind_c == 0
    kind_a, (a1, b1, g1) = S.corrector_rows("ald", sde, sf, ts, 0.16)
    assert kind_a == 1
    for i in (0, 13, N - 1):
        vec_t = torch.ones(3) * ts[i]
        fn = O.rd_predictor_step if pred == "reverse_diffusion" else O.em_predictor_step
        xr, xm = fn(model, spec, x, vec_t, z)
        mine_m = a[i] * x + b[i] * mo
        mine = mine_m + gg[i] * z
        tol = 4e-6 * max(1.0, float(xr.abs().max()))
        assert (mine_m - xm).abs().max() <= tol and (mine - xr).abs().max() <= tol
        # Langevin with batch-mean norms
        xr, xm = O.langevin_step(model, spec, x, vec_t, z, 0.16)
        grad = sc[i] * mo
        gn = grad.reshape(3, -1).norm(dim=-1).mean()
        zn = z.reshape(3, -1).norm(dim=-1).mean()
        step = (0.16 * zn / gn) ** 2 * 2 * al[i]
        mine_m = x + step * grad
        mine = mine_m + torch.sqrt(step * 2) * z
        tol = 4e-6 * max(1.0, float(xr.abs().max()))
        assert (mine_m - xm).abs().max() <= tol and (mine - xr).abs().max() <= tol
        xr, xm = O.ald_step(model, spec, x, vec_t, z, 0.16)
        mine_m = a1[i] * x + b1[i] * mo
        assert (mine_m - xm).abs().max() <= tol and (mine_m + g1[i] * z - xr).abs().max() <= tol
    # time labels fed to the U-Net
    lab = sf.labels(ts)
    if kind == "vp":
        assert torch.equal(lab, ts * (N - 1))
    else:
        assert torch.allclose(lab, spec.marginal_std(ts))


@pytest.mark.parametrize("name", list(CFGS))
def test_unet_state_dict_is_key_compatible(name):
    cfg, _, _ = CFGS[name]
    u = make_unet(cfg)
    want = O.unet_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in u.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    u.load_state_dict(O.random_state_dict(cfg, seed=0), strict=True)
    assert all(not p.requires_grad for p in u.parameters())


def test_plain_tail_conv_bn_act_keys_and_oracle(golden_extra):
    """resnet_block_order='conv_bn_act' (reference modules/unet.py:112-116): final_conv.1 is the bare 1x1; oracle vs executed reference."""
    cfg = dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, order="conv_bn_act")
    u = M.Unet(None, dim=32, dim_mults=[1, 2], channels=3, use_convnext=False, resnet_block_order="conv_bn_act")
    want = {k: tuple(v) for k, v in O.unet_param_shapes(cfg).items()}
    assert {k: tuple(v.shape) for k, v in u.state_dict().items()} == want
    assert want["final_conv.1.weight"] == (3, 32, 1, 1) and "final_conv.3.weight" not in want
    sd = O.random_state_dict(cfg, seed=0)
    u.load_state_dict(sd, strict=True)
    x = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(11))
    y = O.unet_forward(sd, cfg, x, torch.from_numpy(golden_extra["tiny_cba/t"]).float())
    assert torch.equal(y, torch.from_numpy(golden_extra["tiny_cba/y"]))


def test_loop_argument_validation_and_guidance_flag():
    """run_native_loop rejects shapes the U-Net was not built for BEFORE touching the device path (the reference raises a conv shape
    error); guidance is an explicit on/off (None = off, 0.0 = a valid weight)."""
    u = make_unet(CFGS["tiny"][0])          # 1 channel
    s = M.GaussianDiffusion(10, "linear")
    with pytest.raises((ValueError, L.DmnError)):
        R.run_native_loop(u, kind=L.LOOP_DDPM, shape=[2, 3, 16, 16], device="cuda:0", times=torch.zeros(10), coef=torch.zeros(10, 8))
    d = L.LoopDesc()
    assert hasattr(d, "cfg_on") and hasattr(d, "state_elems")
    assert s.guidance_scale is None


def test_unet_ctor_contract():
    with pytest.raises(NotImplementedError):
        M.Unet(None, dim=32)                      # reference default use_convnext=True: not on the built path
    with pytest.raises(ValueError):
        M.Unet(None, dim=32, use_convnext=False, resnet_block_order="bad")
    u = M.Unet(None, dim=32, dim_mults=[1, 2], use_convnext=False, learned_variance=True)
    assert u.out_dim == 6 and u.in_out_list == [(32, 32), (32, 64)]
    with pytest.raises(L.DmnError):
        u(torch.zeros(1, 3, 16, 16), torch.zeros(1))      # CPU tensor: no fallback


def test_no_cpu_fallback_in_samplers():
    u = make_unet(CFGS["tiny"][0])
    for s in (M.GaussianDiffusion(10, "linear"), M.GeneralizedGaussianDiffusion(10, "linear", ddim_timesteps=5)):
        with pytest.raises(L.DmnError):
            s.sample(u, [2, 1, 16, 16], device="cpu")
    pc = M.PredictorCorrectorSampler("reverse_diffusion", "langevin", 0.16)
    pc.update_sde(M.VPSDE(N=10))
    with pytest.raises(L.DmnError):
        pc.sample(u, [2, 1, 16, 16], device="cpu")


def test_model_resolution():
    u = make_unet(CFGS["tiny_cls"][0])
    lab = torch.tensor([1, 2])
    assert R.resolve_model(u) == (u, None)
    un, cl = R.resolve_model(functools.partial(u.forward, classes=lab))
    assert un is u and cl is lab
    un, cl = R.resolve_model(functools.partial(u, classes=lab))
    assert un is u and cl is lab
    assert R.resolve_model(lambda x, t: x) == (None, None)


def test_reference_helper_api_on_cpu():
    """q_sample / q_posterior / predict_start_from_noise keep the reference's torch semantics (training-side callers)."""
    s = M.GaussianDiffusion(50, "linear")
    tb = O.ddpm_tables(50, "linear")
    g = torch.Generator().manual_seed(3)
    x0, noise = torch.randn(4, 3, 8, 8, generator=g), torch.randn(4, 3, 8, 8, generator=g)
    t = torch.tensor([0, 7, 31, 49])
    xt = s.q_sample(x0, t, noise)
    ref = O._ext(tb["sqrt_alphas_cumprod"], t) * x0 + O._ext(tb["sqrt_one_minus_alphas_cumprod"], t) * noise
    assert torch.equal(xt, ref)
    assert torch.allclose(s.predict_start_from_noise(xt, t, noise), x0, atol=2e-4)
    mean, logvar = s.q_posterior(x0, xt, t)
    assert mean.shape == x0.shape and logvar.shape == (4, 1, 1, 1)
    m, v, lv = s.q_mean_variance(x0, t)
    assert torch.equal(lv, O._ext(tb["log_one_minus_alphas_cumprod"], t))
    assert s.extract(s.betas, t, x0.shape).shape == (4, 1, 1, 1)


def test_wavegrad_tables_and_api(golden):
    """WaveGradDiffusion drop-in: extra tables bit-exact vs the executed reference, coefficient column swapped, native Unet refused."""
    import diffusion_model_nemo_b200.modules as M

    s = M.WaveGradDiffusion(1000, "linear")
    g = golden["tables"]
    for k in ("sqrt_alphas_cumprod_prev", "sqrt_alphas_cumprod_m1"):
        assert torch.equal(getattr(s, k), torch.from_numpy(g[f"wavegrad/linear/1000/{k}"])), k
    ts = torch.tensor([999, 3, 0])
    rows = s._step_rows(ts)
    assert torch.equal(rows[1], s.sqrt_alphas_cumprod_m1[ts]) and torch.equal(rows[0], s.sqrt_recip_alphas_cumprod[ts])
    lv = s._model_arg(5, 4, "cpu")
    assert lv.shape == (4, 1, 1, 1) and float(lv[0]) == float(s.sqrt_alphas_cumprod_prev[6])
    x0 = torch.randn(2, 3, 8, 8)
    noise = torch.randn_like(x0)
    lvl = torch.full((2, 1, 1, 1), 0.8)
    assert torch.allclose(s.q_sample(x0, lvl, noise), 0.8 * x0 + (1 - 0.64) ** 0.5 * noise, atol=1e-5)


def test_film_frequency_table_matches_reference_positional_encoding():
    """engine.film_freqs builds the per-column exponents of the FiLM positional encoding with the reference's own torch ops
    (parts/film.py:19-21): 5000 * level * freq must reproduce the oracle's sin / cos arguments bit for bit."""
    import torch
    from diffusion_model_nemo_b200.engine import film_freqs
    from oracle import ref_port as O

    chans = [32, 64, 128]
    f = film_freqs(chans)
    assert f.shape == (sum(chans),)
    level = torch.tensor([0.37]).view(1, 1, 1, 1)
    col = 0
    for c in chans:
        pe = O.film_positional_encoding(level, c).flatten()
        arg = (5000 * level.view(1) * f[col:col + c])                      # same left-to-right product as the reference
        want = torch.cat([arg[:c // 2].sin(), arg[c // 2:].cos()])
        assert torch.equal(pe, want), c
        assert torch.equal(f[col:col + c // 2], f[col + c // 2:col + c])     # [exponents | exponents]
        col += c


def test_bpd_prior_term_is_schedule_only():
    """The prior term of the bits-per-dimension evaluation depends on x_0 and the schedule alone (abstract_diffusion_model.py:180-184)."""
    import torch
    from oracle import ref_port as O

    tb = O.ddpm_tables(50, "linear")
    x0 = torch.rand(3, 1, 8, 8) * 2 - 1
    a = O.bits_per_dimension(lambda x, t: torch.zeros_like(x), x0, tb, O.NoiseQueue(1))
    b = O.bits_per_dimension(lambda x, t: torch.ones_like(x), x0, tb, O.NoiseQueue(2))
    assert torch.equal(a["prior_bpd"], b["prior_bpd"]) and (a["prior_bpd"] > 0).all()
    assert a["terms_bpd"].shape == (3, 50) and not torch.equal(a["terms_bpd"], b["terms_bpd"])
