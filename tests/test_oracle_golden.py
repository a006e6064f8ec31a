"""The CPU oracle (oracle/ref_port.py) against the fixtures produced by EXECUTING the reference
(tests/golden/make_golden.py).  This is the pin of the oracle (the reference ships no tests of its own)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import CFGS, GOLDEN
from oracle import ref_port as O


def _same_cpu_path():
    meta = json.load(open(os.path.join(GOLDEN, "golden_meta.json")))["meta"]
    return meta["cpu_capability"] == torch.backends.cpu.get_cpu_capability() and meta["torch"] == torch.__version__


def assert_table(mine: torch.Tensor, gold: np.ndarray, what):
    gold = torch.from_numpy(gold)
    if torch.equal(mine, gold):
        return
    # a different host ISA may pick another vectorised libm path: allow <= 4 ULP there, never on the fixture host
    assert not _same_cpu_path(), f"{what}: not bit-exact on the fixture's own CPU path"
    ulp = (mine.view(torch.int32).long() - gold.view(torch.int32).long()).abs().max()
    assert ulp <= 4, f"{what}: {int(ulp)} ULP"


@pytest.mark.parametrize("name", ["linear", "quadratic", "sigmoid", "cosine"])
@pytest.mark.parametrize("T", [50, 250, 1000])
def test_ddpm_tables_bit_exact(golden, name, T):
    mine = O.ddpm_tables(T, name)
    for k in O.DDPM_TABLE_NAMES:
        assert_table(mine[k], golden["tables"][f"ddpm/{name}/{T}/{k}"], (name, T, k))


def test_custom_schedule_cfg(golden):
    mine = O.ddpm_tables(100, "linear", {"linear": {"beta_start": 1e-3, "beta_end": 0.05}})
    for k in O.DDPM_TABLE_NAMES:
        assert_table(mine[k], golden["tables"][f"ddpm_custom/linear/100/{k}"], k)


def test_ddim_vp_ve_wavegrad_tables(golden):
    g = golden["tables"]
    for name, T, S in (("cosine", 1000, 50), ("linear", 1000, 10), ("linear", 20, 5)):
        assert_table(O.ddim_extended_cumprod(O.ddpm_tables(T, name)["betas"]), g[f"ddim/{name}/{T}/alphas_extended_cumprod"], name)
        assert O.ddim_pairs(T, S) == [tuple(p) for p in g[f"ddim/{name}/{T}/{S}/pairs"].tolist()]
    assert O.ddim_pairs(1000, 50)[0] == (980, 960) and O.ddim_pairs(1000, 50)[-1] == (0, -1)
    wg = O.wavegrad_tables(O.ddpm_tables(1000, "linear"))
    for k in ("sqrt_alphas_cumprod_prev", "sqrt_alphas_cumprod_m1"):
        assert_table(wg[k], g[f"wavegrad/linear/1000/{k}"], k)
    vp, ve = O.vp_tables(0.1, 20.0, 1000), O.ve_tables(0.01, 50.0, 1000)
    for k, v in vp.items():
        assert_table(v, g[f"vp/1000/{k}"], k)
    assert_table(ve["discrete_sigmas"], g["ve/1000/discrete_sigmas"], "ve")


@pytest.mark.parametrize("name", list(CFGS))
def test_unet_forward_matches_reference(golden, name):
    cfg, size, b = CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    x = torch.from_numpy(golden["unet"][f"{name}/x"])
    # the fixture input itself is regenerated from its seed: guards against generator drift
    assert torch.equal(x, torch.randn(b, cfg["channels"], size, size, generator=torch.Generator().manual_seed(11)))
    cls = torch.tensor([3, 10][:b]) if cfg.get("num_classes") is not None else None
    for tname in ("int", "float"):
        t = torch.from_numpy(golden["unet"][f"{name}/{tname}/t"])
        y = O.unet_forward(sd, cfg, x, t, cls)
        ref = torch.from_numpy(golden["unet"][f"{name}/{tname}/y"])
        assert (y - ref).abs().max() <= 2e-5 * ref.abs().max(), (name, tname)


def test_flop_counter_matches_torch_counter():
    meta = json.load(open(os.path.join(GOLDEN, "golden_meta.json")))
    for name, (cfg, size, _) in CFGS.items():
        assert O.unet_flops_per_sample(cfg, size) == meta[f"unet/{name}/flops_per_sample"]
    assert O.unet_flops_per_sample(CFGS["cfg2"][0], 32) == 5350096896
    assert O.unet_flops_per_sample(CFGS["cfg1"][0], 28) == 412641792


def test_samplers_match_reference(golden):
    g = golden["samplers"]
    cfg, size, b = CFGS["tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    model = O.make_model(sd, cfg)
    shape = [b, cfg["channels"], size, size]
    for sched in ("linear", "cosine"):
        final, _ = O.sample_ddpm(model, shape, O.ddpm_tables(20, sched), O.NoiseQueue(5))
        assert ((final + 1) * 0.5 - torch.from_numpy(g[f"ddpm/{sched}/20/final01"])).abs().max() < 2e-5
    for eta in (0.0, 0.5):
        final = O.sample_ddim(model, shape, O.ddpm_tables(20, "linear"), O.NoiseQueue(5), eta=eta, ddim_timesteps=5)
        assert ((final + 1) * 0.5 - torch.from_numpy(g[f"ddim/linear/20/5/eta{eta}/final01"])).abs().max() < 2e-5
    cfg_lv = CFGS["tiny_lv"][0]
    sd_lv = O.random_state_dict(cfg_lv, seed=0)
    final, _ = O.sample_ddpm(O.make_model(sd_lv, cfg_lv), [b, 3, size, size], O.ddpm_tables(20, "cosine"), O.NoiseQueue(5), kind="learned")
    assert ((final + 1) * 0.5 - torch.from_numpy(g["learned/cosine/20/final01"])).abs().max() < 2e-5


@pytest.mark.parametrize("kind", ["vp", "ve"])
@pytest.mark.parametrize("pc", [("reverse_diffusion", "langevin"), ("euler_maruyama", "none"), ("reverse_diffusion", "ald")])
def test_pc_sampler_matches_reference(golden, kind, pc):
    cfg, size, b = CFGS["tiny_g4"]
    sd = O.random_state_dict(cfg, seed=0)
    model = O.make_model(sd, cfg)
    for dn in (1, 0):
        last, _ = O.sample_pc(model, [b, 3, size, size], O.SDESpec(kind, N=40), O.NoiseQueue(5), predictor=pc[0], corrector=pc[1],
                              snr=0.16, n_steps=1, denoise=bool(dn))
        ref = torch.from_numpy(golden["samplers"][f"pc/{kind}/{pc[0]}/{pc[1]}/dn{dn}/final01"]) * 2 - 1
        assert (last - ref).abs().max() <= 5e-5 * max(1.0, float(ref.abs().max()))


def test_teacher_forced_step_fixture(golden):
    cfg, size, b = CFGS["cfg2"]
    sd = O.random_state_dict(cfg, seed=0)
    x = torch.from_numpy(golden["step"]["cfg2/x"])
    tb = O.ddpm_tables(1000, "linear")
    ti = 500
    t = torch.full((b,), ti, dtype=torch.long)
    eps = O.unet_forward(sd, cfg, x, t.float())
    ref = torch.from_numpy(golden["step"][f"cfg2/t{ti}/eps"])
    assert (eps - ref).abs().max() <= 2e-5 * ref.abs().max()
    xn = O.ddpm_step(tb, x, t, eps, O.NoiseQueue(9)(x.shape))
    assert (xn - torch.from_numpy(golden["step"][f"cfg2/t{ti}/x_next"])).abs().max() < 2e-5


def test_wavegrad_sampler_oracle_vs_executed_reference(golden):
    """WaveGradDiffusion (reference modules/wavegrad_diffusion.py) with the stand-in denoiser: the oracle loop reproduces the
    fixture written by executing the reference (tests/golden/make_golden_wavegrad.py)."""
    shape = [2, 3, 16, 16]
    for sched, T in (("linear", 20), ("cosine", 50)):
        mine = O.sample_wavegrad(O.wavegrad_toy_model, shape, O.ddpm_tables(T, sched), O.NoiseQueue(5))
        ref01 = torch.from_numpy(golden["wavegrad"][f"wavegrad/{sched}/{T}/final01"])
        assert ((mine + 1) * 0.5 - ref01).abs().max() <= 2e-5
