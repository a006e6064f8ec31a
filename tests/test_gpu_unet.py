"""U-Net forward parity on the GPU against the golden outputs of the executed reference (tests/golden/unet.npz,
step.npz): rel-L2 <= 1e-4 in fp32 mode, <= 2e-2 in bf16 mode (BASELINE.json north_star tolerances)."""
import pytest
import torch

from conftest import CFGS, make_unet, rel_l2
from oracle import ref_port as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"fp32": 1e-4, "bf16": 2e-2}


def _run(name, golden, dtype, engine):
    cfg, size, b = CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype=dtype, engine=engine, device=DEV)
    x = torch.from_numpy(golden["unet"][f"{name}/x"]).to(DEV)
    cls = torch.tensor([3, 10][:b], device=DEV) if cfg.get("num_classes") is not None else None
    errs = {}
    for tname in ("int", "float"):
        t = torch.from_numpy(golden["unet"][f"{name}/{tname}/t"]).to(DEV)
        y = u(x, t, cls) if cls is not None else u(x, t)
        ref = torch.from_numpy(golden["unet"][f"{name}/{tname}/y"])
        assert y.shape == ref.shape and torch.isfinite(y).all()
        errs[tname] = rel_l2(y.cpu(), ref)
    return errs


@pytest.mark.parametrize("name", list(CFGS))
def test_unet_fp32_vs_reference_golden(golden, name):
    errs = _run(name, golden, "fp32", "simt")
    assert max(errs.values()) <= TOL["fp32"], errs


@pytest.mark.parametrize("name", list(CFGS))
def test_unet_bf16_simt_vs_reference_golden(golden, name):
    errs = _run(name, golden, "bf16", "simt")
    assert max(errs.values()) <= TOL["bf16"], errs


def test_unet_int_time_and_partial_batch(golden):
    """time may be int64 (DDPM) or float (score-SDE); a plan sized for a larger batch serves smaller ones."""
    cfg, size, b = CFGS["tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="fp32", engine="simt", device=DEV)
    x = torch.from_numpy(golden["unet"]["tiny/x"]).to(DEV)
    t = torch.tensor([7, 513], device=DEV)                      # int64, as GaussianDiffusion passes it
    big = u(torch.cat([x, x, x]), torch.cat([t, t, t]))
    small = u(x, t)                                               # same plan (max_batch 6), batch 2
    ref = torch.from_numpy(golden["unet"]["tiny/int/y"])
    assert rel_l2(small.cpu(), ref) <= 1e-4 and rel_l2(big[4:].cpu(), ref) <= 1e-4
    assert len(u._plans) == 1
    # in-place weight updates are picked up (parameter version tracking)
    with torch.no_grad():
        u.final_conv[3].bias.add_(1.0)
    assert torch.allclose(u(x, t), small + 1.0, atol=1e-5)


def test_teacher_forced_eps_cfg2(golden):
    """Per-step eps on the reference's own x_t (teacher forced), CIFAR-shape U-Net, t in {999, 500, 1, 0}."""
    cfg, size, b = CFGS["cfg2"]
    sd = O.random_state_dict(cfg, seed=0)
    x = torch.from_numpy(golden["step"]["cfg2/x"]).to(DEV)
    for dtype, engine in (("fp32", "simt"), ("bf16", "simt")):
        u = make_unet(cfg, sd, dtype=dtype, engine=engine, device=DEV)
        for ti in (999, 500, 1, 0):
            eps = u(x, torch.full((b,), ti, device=DEV))
            ref = torch.from_numpy(golden["step"][f"cfg2/t{ti}/eps"])
            assert rel_l2(eps.cpu(), ref) <= TOL[dtype], (dtype, ti)
