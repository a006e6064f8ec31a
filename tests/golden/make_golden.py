"""Generate the golden fixtures under tests/golden/ by EXECUTING the unmodified reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does
  1. imports titu1994/diffusion_model_nemo from /root/reference through the stub shim in
     tests/_shim (nemo / pytorch_lightning / omegaconf / hydra are absent offline);
  2. asserts that the CPU oracle (oracle/ref_port.py) reproduces the reference: schedule tables
     bit-exact, U-Net forward and every sampler to fp32 round-off (the same aten ops in the same order
     usually give 0 difference);
  3. writes small fixtures (inputs are regenerated from seeds, outputs are stored) that the CPU
     test-suite and the GPU parity tests re-check.

The reference has no tests or golden vectors of its own: these files ARE the pin.
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests", "_shim"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

import diffusion_model_nemo.modules as M  # noqa: E402  (the unmodified reference)
from oracle import ref_port as O  # noqa: E402

torch.set_num_threads(8)

CFGS = {
    # name: (unet cfg, image size, batch)
    "tiny": (dict(dim=32, dim_mults=[1, 2], channels=1, groups=8), 16, 2),
    "cfg1": (dict(dim=32, dim_mults=[1, 2, 4], channels=1, groups=8), 28, 2),
    "cfg2": (dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8), 32, 1),
    "tiny_lv": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, learned_variance=True), 16, 2),
    "tiny_g4": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=4), 16, 2),
    "tiny_cls": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, num_classes=10), 16, 2),
}


def build_ref_unet(cfg, sd):
    u = M.Unet(input_dim=None, dim=cfg["dim"], dim_mults=cfg["dim_mults"], channels=cfg["channels"],
               use_convnext=False, resnet_block_groups=cfg["groups"], dropout=0.0,
               learned_variance=cfg.get("learned_variance", False), num_classes=cfg.get("num_classes")).eval()
    missing = u.load_state_dict(sd, strict=True)       # key-for-key: proves the parameter naming of the port
    assert not missing.missing_keys and not missing.unexpected_keys
    return u


@contextlib.contextmanager
def injected_noise(queue):
    """Swap torch.randn / torch.randn_like for the FIFO (reference draw order)."""
    orig_randn, orig_like = torch.randn, torch.randn_like

    def randn(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple, torch.Size)):
            shape = tuple(shape[0])
        return queue(shape)

    def randn_like(x, **kw):
        return queue(tuple(x.shape))

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like


def maxdiff(a, b):
    return float((a - b).abs().max())


def main():
    out = {}
    meta = {"torch": torch.__version__, "cpu_capability": torch.backends.cpu.get_cpu_capability()}

    # ---- 1. schedule tables --------------------------------------------------------------------
    tables = {}
    for name in ("linear", "quadratic", "sigmoid", "cosine"):
        for T in (50, 250, 1000):
            ref = M.GaussianDiffusion(timesteps=T, schedule_name=name)
            mine = O.ddpm_tables(T, name)
            for k in O.DDPM_TABLE_NAMES:
                assert torch.equal(getattr(ref, k), mine[k]), (name, T, k)
                tables[f"ddpm/{name}/{T}/{k}"] = mine[k].numpy()
    # custom schedule cfg (schedule_cfg is a dict keyed by schedule name, gaussian_diffusion.py:56-58)
    ref = M.GaussianDiffusion(timesteps=100, schedule_name="linear", schedule_cfg={"linear": {"beta_start": 1e-3, "beta_end": 0.05}})
    mine = O.ddpm_tables(100, "linear", {"linear": {"beta_start": 1e-3, "beta_end": 0.05}})
    for k in O.DDPM_TABLE_NAMES:
        assert torch.equal(getattr(ref, k), mine[k])
        tables[f"ddpm_custom/linear/100/{k}"] = mine[k].numpy()
    # DDIM extended table + index pairs
    for name, T, S in (("cosine", 1000, 50), ("linear", 1000, 10), ("linear", 20, 5)):
        ref = M.GeneralizedGaussianDiffusion(timesteps=T, schedule_name=name, eta=0.0, ddim_timesteps=S)
        ref.betas_extended = torch.cat([torch.zeros(1), ref.betas], dim=0)
        aext = (1.0 - ref.betas_extended).cumprod(dim=0)
        assert torch.equal(aext, O.ddim_extended_cumprod(O.ddpm_tables(T, name)["betas"]))
        tables[f"ddim/{name}/{T}/alphas_extended_cumprod"] = aext.numpy()
        stride = T // S
        seq = list(range(0, T, stride))
        pairs = list(zip(reversed(seq), reversed([-1] + seq[:-1])))
        assert pairs == O.ddim_pairs(T, S)
        tables[f"ddim/{name}/{T}/{S}/pairs"] = np.asarray(pairs, dtype=np.int64)
    # WaveGrad extras
    ref = M.WaveGradDiffusion(timesteps=1000, schedule_name="linear")
    mine = O.wavegrad_tables(O.ddpm_tables(1000, "linear"))
    for k in ("sqrt_alphas_cumprod_prev", "sqrt_alphas_cumprod_m1"):
        assert torch.equal(getattr(ref, k), mine[k])
        tables[f"wavegrad/linear/1000/{k}"] = mine[k].numpy()
    # VP / VE
    vp, ve = M.VPSDE(0.1, 20.0, 1000), M.VESDE(0.01, 50.0, 1000)
    mvp, mve = O.vp_tables(0.1, 20.0, 1000), O.ve_tables(0.01, 50.0, 1000)
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_1m_alphas_cumprod"):
        assert torch.equal(getattr(vp, k), mvp[k])
        tables[f"vp/1000/{k}"] = mvp[k].numpy()
    assert torch.equal(ve.discrete_sigmas, mve["discrete_sigmas"])
    tables["ve/1000/discrete_sigmas"] = mve["discrete_sigmas"].numpy()
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tables)

    # ---- 2. U-Net forward ----------------------------------------------------------------------
    unet = {}
    for name, (cfg, size, b) in CFGS.items():
        sd = O.random_state_dict(cfg, seed=0)
        ref = build_ref_unet(cfg, sd)
        g = torch.Generator().manual_seed(11)
        x = torch.randn(b, cfg["channels"], size, size, generator=g)
        for tname, t in (("int", torch.tensor([7, 513][:b])), ("float", torch.tensor([0.37 * 999, 12.5][:b]))):
            kw = {}
            if cfg.get("num_classes") is not None:
                kw["classes"] = torch.tensor([3, 10][:b])
            with torch.no_grad():
                y_ref = ref(x, t, **kw)
                y_mine = O.unet_forward(sd, cfg, x, t.float(), kw.get("classes"))
            d = maxdiff(y_ref, y_mine)
            assert d <= 1e-5 * float(y_ref.abs().max()), (name, d)
            unet[f"{name}/{tname}/t"] = t.float().numpy()
            unet[f"{name}/{tname}/y"] = y_ref.numpy()
            out[f"unet/{name}/{tname}/oracle_maxdiff"] = d
        unet[f"{name}/x"] = x.numpy()
        # flop count cross-check vs torch's own counter on the reference module
        from torch.utils.flop_counter import FlopCounterMode
        with FlopCounterMode(display=False) as fc, torch.no_grad():
            ref(x[:1], torch.tensor([5]))
        out[f"unet/{name}/flops_per_sample"] = int(fc.get_total_flops())
        assert int(fc.get_total_flops()) == O.unet_flops_per_sample(cfg, size), (name, fc.get_total_flops(), O.unet_flops_per_sample(cfg, size))
    np.savez_compressed(os.path.join(HERE, "unet.npz"), **unet)

    # ---- 3. samplers (injected noise) ------------------------------------------------------------
    samp = {}
    cfg, size, b = CFGS["tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    ref_unet = build_ref_unet(cfg, sd)
    shape = [b, cfg["channels"], size, size]
    port_model = O.make_model(sd, cfg)

    def run_ref(sampler, model=ref_unet, seed=5, **kw):
        with injected_noise(O.NoiseQueue(seed)), torch.no_grad():
            imgs = sampler.sample(model, shape, device=torch.device("cpu"), **kw)
        return imgs

    # DDPM ancestral, several schedules, T=20
    for sched in ("linear", "cosine"):
        s = M.GaussianDiffusion(timesteps=20, schedule_name=sched)
        imgs = run_ref(s)
        mine, _ = O.sample_ddpm(port_model, shape, O.ddpm_tables(20, sched), O.NoiseQueue(5))
        ref_final = imgs[-1] * 2 - 1
        assert len(imgs) == 20
        assert maxdiff(ref_final, mine) < 2e-5, maxdiff(ref_final, mine)
        samp[f"ddpm/{sched}/20/final01"] = imgs[-1].numpy()      # as returned by the reference: [0,1]
        samp[f"ddpm/{sched}/20/step10_01"] = imgs[9].numpy()
    # learned variance
    cfg_lv, _, _ = CFGS["tiny_lv"]
    sd_lv = O.random_state_dict(cfg_lv, seed=0)
    ref_lv = build_ref_unet(cfg_lv, sd_lv)
    shape3 = [b, 3, size, size]
    s = M.LearnedGaussianDiffusion(timesteps=20, schedule_name="cosine")
    with injected_noise(O.NoiseQueue(5)), torch.no_grad():
        imgs = s.sample(ref_lv, shape3, device=torch.device("cpu"))
    mine, _ = O.sample_ddpm(O.make_model(sd_lv, cfg_lv), shape3, O.ddpm_tables(20, "cosine"), O.NoiseQueue(5), kind="learned")
    assert maxdiff(imgs[-1] * 2 - 1, mine) < 2e-5
    samp["learned/cosine/20/final01"] = imgs[-1].numpy()
    # DDIM
    for eta in (0.0, 0.5):
        s = M.GeneralizedGaussianDiffusion(timesteps=20, schedule_name="linear", eta=eta, ddim_timesteps=5)
        imgs = run_ref(s)
        mine = O.sample_ddim(port_model, shape, O.ddpm_tables(20, "linear"), O.NoiseQueue(5), eta=eta, ddim_timesteps=5)
        assert len(imgs) == 5
        assert maxdiff(imgs[-1] * 2 - 1, mine) < 2e-5
        samp[f"ddim/linear/20/5/eta{eta}/final01"] = imgs[-1].numpy()
    # score-SDE PC (3-channel, groups=4 U-Net as in configs/score_sde/vp/unet_small.yaml:35)
    cfg4, _, _ = CFGS["tiny_g4"]
    sd4 = O.random_state_dict(cfg4, seed=0)
    ref4 = build_ref_unet(cfg4, sd4)
    port4 = O.make_model(sd4, cfg4)
    for kind, sde_ref in (("vp", M.VPSDE(0.1, 20.0, 40)), ("ve", M.VESDE(0.01, 50.0, 40))):
        for pred, corr in (("reverse_diffusion", "langevin"), ("euler_maruyama", "none"), ("reverse_diffusion", "ald")):
            for denoise in (True, False):
                s = M.PredictorCorrectorSampler(pred, corr, snr=0.16, n_steps=1, denoise=denoise)
                s.update_sde(sde_ref)
                with injected_noise(O.NoiseQueue(5)), torch.no_grad():
                    imgs = s.sample(ref4, shape3, device=torch.device("cpu"))
                spec = O.SDESpec(kind, N=40)
                last, _ = O.sample_pc(port4, shape3, spec, O.NoiseQueue(5), predictor=pred, corrector=corr,
                                      snr=0.16, n_steps=1, denoise=denoise)
                ref_final = imgs[-1] * 2 - 1
                d = maxdiff(ref_final, last)
                assert d <= 2e-5 * max(1.0, float(ref_final.abs().max())), (kind, pred, corr, d)
                samp[f"pc/{kind}/{pred}/{corr}/dn{int(denoise)}/final01"] = imgs[-1].numpy()
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), **samp)

    # ---- 4. teacher-forced step fixture on cfg2 (one p_sample step, stored eps + x_{t-1}) ----------
    cfg, size, b = CFGS["cfg2"]
    sd = O.random_state_dict(cfg, seed=0)
    ref = build_ref_unet(cfg, sd)
    s = M.GaussianDiffusion(timesteps=1000, schedule_name="linear")
    g = torch.Generator().manual_seed(21)
    x = torch.randn(b, 3, size, size, generator=g)
    step = {}
    for ti in (999, 500, 1, 0):
        t = torch.full((b,), ti, dtype=torch.long)
        with injected_noise(O.NoiseQueue(9)), torch.no_grad():
            eps = ref(x, t)
            xn = s.p_sample(ref, x, t)
        z = O.NoiseQueue(9)(x.shape)
        mine = O.ddpm_step(O.ddpm_tables(1000, "linear"), x, t, O.unet_forward(sd, cfg, x, t.float()), z)
        assert maxdiff(xn, mine) < 2e-5
        step[f"cfg2/t{ti}/eps"] = eps.numpy()
        step[f"cfg2/t{ti}/x_next"] = xn.numpy()
    step["cfg2/x"] = x.numpy()
    np.savez_compressed(os.path.join(HERE, "step.npz"), **step)

    out["meta"] = meta
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
    for fn in ("tables.npz", "unet.npz", "samplers.npz", "step.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
