"""Pins of the PRODUCTION engine (bf16 activations, tcgen05 convolutions, fused attention) at the benchmark shapes.

  (a) BASELINE configs[1] teacher-forced eps against the CPU oracle at batch 64 (mixed 256/128-row tiling, multi-round persistent
      CTAs, every instantiation the batch-256 benchmark launches), t in {999, 500, 1, 0}; batch-256 vs batch-64 slice equality in
      GRAPH mode (in-kernel Philox is keyed by the element index, so the first 64 samples see the same noise in both runs).
  (b) class-conditional U-Net on the tensor-core engine against fixtures of the executed reference (stem epilogue adds the embedding).
  (c) configs[2] / configs[3] at their real shapes: per-evaluation model output along the ORACLE trajectory (teacher forced) for 3 steps.
  (d) bf16 free-running drift: final-sample max-abs against the fp32 CPU oracle with the same weights and the same injected noise;
      gates are 2x the values measured on a B200.

Tolerances: per-evaluation rel-L2 <= 2e-2 (north_star, bf16 mode).
"""
import pytest
import torch

import diffusion_model_nemo_b200.modules as M
from conftest import CFGS, make_unet, rel_l2
from oracle import ref_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG2 = dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8)


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


class Recorder:
    """Wraps an oracle model callable and records every (x, t, out) it sees."""

    def __init__(self, fn):
        self.fn, self.calls = fn, []

    def __call__(self, x, t, *a, **k):
        out = self.fn(x, t, *a, **k)
        self.calls.append((x.clone(), t.clone(), out.clone()))
        return out


def test_cfg2_teacher_forced_eps_batch64_vs_oracle():
    sd = O.random_state_dict(CFG2, seed=0)
    u = make_unet(CFG2, sd, dtype="bf16", engine="tcgen05", device=DEV)
    x = _rand(64, 3, 32, 32, seed=21)
    t = torch.tensor([999, 500, 1, 0]).repeat_interleave(16)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        ref = O.unet_forward(sd, CFG2, x, t.float())
    eps = u(x.to(DEV), t.to(DEV)).cpu()
    for k, ti in enumerate((999, 500, 1, 0)):
        sl = slice(16 * k, 16 * k + 16)
        assert rel_l2(eps[sl], ref[sl]) <= 2e-2, ti
    # per-sample: no sample may be far off while the batch average looks fine
    per = ((eps - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1))
    assert float(per.max()) <= 3e-2


def test_cfg2_batch4_vs_reference_golden(golden_extra):
    sd = O.random_state_dict(CFG2, seed=0)
    u = make_unet(CFG2, sd, dtype="bf16", engine="tcgen05", device=DEV)
    x = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(11))
    t = torch.from_numpy(golden_extra["cfg2_b4/t"])
    y = u(x.to(DEV), t.to(DEV)).cpu()
    assert rel_l2(y, torch.from_numpy(golden_extra["cfg2_b4/y"])) <= 2e-2


def test_cfg2_graph_mode_batch256_equals_batch64_slice():
    u = make_unet(CFG2, O.random_state_dict(CFG2, seed=0), dtype="bf16", engine="tcgen05", device=DEV)
    s = M.GaussianDiffusion(8, "linear")
    s.seed = 4242
    a = s.sample(u, [256, 3, 32, 32], device=DEV)[-1]
    a2 = s.sample(u, [256, 3, 32, 32], device=DEV)[-1]
    c = s.sample(u, [64, 3, 32, 32], device=DEV)[-1]
    assert torch.isfinite(a).all() and torch.equal(a, a2)
    assert (a[:64] - c).abs().max() <= 1e-6
    # graph replay == plain launches at the benchmark batch
    s.use_cuda_graph = False
    p = s.sample(u, [256, 3, 32, 32], device=DEV)[-1]
    assert torch.equal(a, p)


@pytest.mark.parametrize("name", ["tiny_cls", "cfg2_cls"])
def test_class_conditional_unet_on_tcgen05_vs_reference_golden(golden, golden_extra, name):
    if name == "tiny_cls":
        cfg, size, b = CFGS[name]
        x = torch.from_numpy(golden["unet"][f"{name}/x"])
        t = torch.from_numpy(golden["unet"][f"{name}/int/t"])
        y_ref = torch.from_numpy(golden["unet"][f"{name}/int/y"])
        classes = torch.tensor([3, 10][:b])
    else:
        cfg = dict(CFG2, num_classes=10)
        x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(11))
        t = torch.from_numpy(golden_extra["cfg2_cls/t"])
        y_ref = torch.from_numpy(golden_extra["cfg2_cls/y"])
        classes = torch.from_numpy(golden_extra["cfg2_cls/classes"])
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    y = u(x.to(DEV), t.to(DEV), classes.to(DEV)).cpu()
    assert rel_l2(y, y_ref) <= 2e-2
    # the label matters, and classes=None means the padding row (reference unet.py:135-137)
    y_null = u(x.to(DEV), t.to(DEV)).cpu()
    ref_null = O.unet_forward(sd, cfg, x, t.float(), None)
    assert rel_l2(y_null, ref_null) <= 2e-2
    assert rel_l2(y, ref_null) > rel_l2(y, y_ref)


def test_plain_tail_conv_bn_act_vs_reference_golden(golden_extra):
    """resnet_block_order='conv_bn_act': final_conv = [ResnetBlock, Conv2d 1x1] (reference modules/unet.py:112-116)."""
    cfg = dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, order="conv_bn_act")
    sd = O.random_state_dict(cfg, seed=0)
    x = torch.randn(2, 3, 16, 16, generator=torch.Generator().manual_seed(11))
    t = torch.from_numpy(golden_extra["tiny_cba/t"])
    y_ref = torch.from_numpy(golden_extra["tiny_cba/y"])
    for dtype, engine, tol in (("fp32", "simt", 1e-4), ("bf16", "tcgen05", 2e-2)):
        u = M.Unet(None, dim=32, dim_mults=[1, 2], channels=3, use_convnext=False, resnet_block_groups=8, resnet_block_order="conv_bn_act",
                   compute_dtype=dtype, conv_engine=engine)
        u.load_state_dict(sd, strict=True)
        u.to(DEV)
        assert rel_l2(u(x.to(DEV), t.to(DEV)).cpu(), y_ref) <= tol, (dtype, engine)


def test_config3_teacher_forced_along_oracle_trajectory():
    """configs[2]: 3x64x64 learned-variance U-Net, 250-step cosine table; model output at the oracle's x_t for the first 3 steps."""
    cfg = dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8, learned_variance=True)
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    tb = O.ddpm_tables(250, "cosine")
    q = O.NoiseQueue(5)
    shape = [2, 3, 64, 64]
    rec = Recorder(O.make_model(sd, cfg))
    img = q(shape)
    with torch.no_grad():
        for i in (249, 248, 247):
            t = torch.full((2,), i, dtype=torch.long)
            img = O.learned_step(tb, img, t, rec(img, t), q(shape))
    assert len(rec.calls) == 3
    for x, t, out in rec.calls:
        y = u(x.to(DEV), t.to(DEV)).cpu()
        assert y.shape == out.shape == (2, 6, 64, 64)
        assert rel_l2(y, out) <= 2e-2, int(t[0])


@pytest.mark.parametrize("kind", ["vp", "ve"])
def test_config4_teacher_forced_scores_along_oracle_trajectory(kind):
    """configs[3]: reverse-diffusion predictor + Langevin corrector, groups = 4; the denoiser output of every evaluation (2 per
    step) along the oracle's trajectory for 3 steps (VE states reach |x| ~ 50 and the time input is sigma)."""
    cfg = dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=4)
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    sde = O.SDESpec(kind, N=1000)
    rec = Recorder(O.make_model(sd, cfg))
    with torch.no_grad():
        O.sample_pc(rec, [2, 3, 32, 32], sde, O.NoiseQueue(7), n_iter=3)
    assert len(rec.calls) == 6
    for x, t, out in rec.calls:
        y = u(x.to(DEV), t.to(DEV)).cpu()
        assert rel_l2(y, out) <= 2e-2, (kind, float(t[0]))


# ---- (d) bf16 free-running drift -------------------------------------------------------------------------------------------------
# Measured on a B200 (round 2, batch 2, seeded random-init weights, injected noise; max-abs on the [0,1] image):
#   ddpm1000 1.48e-2   (ancestral sampling re-injects noise every step and clamps x0: bf16 error does not accumulate)
#   ddim50   3.49e-1   (eta = 0: a deterministic map through an UNTRAINED network is chaotic; per-step eps error stays <= 2e-2)
#   pc40     1.76      (VP predictor-corrector, unclamped state; same chaos, Langevin step size couples the batch)
# Gates: the SURVEY section 8(d) proposal for the headline run (5e-2) and 2x the measured value for the two chaotic loops.
DRIFT_GATES = {"ddpm1000": 5e-2, "ddim50": 0.70, "pc40": 3.6}


def _drift_report(name, d):
    print(f"\n[bf16 drift] {name}: final-sample max-abs vs fp32 oracle = {d:.4e}")
    gate = DRIFT_GATES[name]
    if gate is not None:
        assert d <= gate, (name, d, gate)
    assert d == d


def test_bf16_drift_ddim50_vs_fp32_oracle():
    sd = O.random_state_dict(CFG2, seed=0)
    u = make_unet(CFG2, sd, dtype="bf16", engine="tcgen05", device=DEV)
    shape = [2, 3, 32, 32]
    q = O.NoiseQueue(5)
    noise = torch.stack([q(shape) for _ in range(51)])
    d = M.GeneralizedGaussianDiffusion(1000, "linear", eta=0.0, ddim_timesteps=50)
    got = d.sample(u, shape, device=DEV, noise=noise)[-1]
    with torch.no_grad():
        ref = O.sample_ddim(O.make_model(sd, CFG2), shape, O.ddpm_tables(1000, "linear"), O.NoiseQueue(5), eta=0.0, ddim_timesteps=50)
    _drift_report("ddim50", float((got - (ref + 1) * 0.5).abs().max()))


def test_bf16_drift_pc40_vs_fp32_oracle():
    cfg = dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=4)
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    shape = [2, 3, 32, 32]
    q = O.NoiseQueue(5)
    noise = torch.stack([q(shape) for _ in range(81)])
    s = M.PredictorCorrectorSampler("reverse_diffusion", "langevin", snr=0.16, n_steps=1, denoise=True)
    s.update_sde(M.VPSDE(0.1, 20.0, 40))
    got = s.sample(u, shape, device=DEV, noise=noise)[-1]
    with torch.no_grad():
        ref, _ = O.sample_pc(O.make_model(sd, cfg), shape, O.SDESpec("vp", N=40), O.NoiseQueue(5))
    _drift_report("pc40", float((got - (ref + 1) * 0.5).abs().max()))


def test_bf16_drift_ddpm1000_vs_fp32_oracle():
    sd = O.random_state_dict(CFG2, seed=0)
    u = make_unet(CFG2, sd, dtype="bf16", engine="tcgen05", device=DEV)
    shape = [2, 3, 32, 32]
    q = O.NoiseQueue(5)
    noise = torch.stack([q(shape) for _ in range(1001)])
    s = M.GaussianDiffusion(1000, "linear")
    got = s.sample(u, shape, device=DEV, noise=noise)[-1]
    with torch.no_grad():
        ref, _ = O.sample_ddpm(O.make_model(sd, CFG2), shape, O.ddpm_tables(1000, "linear"), O.NoiseQueue(5))
    _drift_report("ddpm1000", float((got - (ref + 1) * 0.5).abs().max()))
