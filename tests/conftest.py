import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("tables", "unet", "samplers", "step", "wavegrad", "wavegrad_unet", "bpd"):
        out[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return out


@pytest.fixture(scope="session")
def golden_extra():
    """Round-2 fixtures of the executed reference (tests/golden/make_golden_extra.py)."""
    return dict(np.load(os.path.join(GOLDEN, "unet_extra.npz")))


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """The native library is the product; build it in-tree if the .so is stale or missing."""
    from diffusion_model_nemo_b200 import _build

    _build.build()


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


CFGS = {
    "tiny": (dict(dim=32, dim_mults=[1, 2], channels=1, groups=8), 16, 2),
    "cfg1": (dict(dim=32, dim_mults=[1, 2, 4], channels=1, groups=8), 28, 2),
    "cfg2": (dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8), 32, 1),
    "tiny_lv": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, learned_variance=True), 16, 2),
    "tiny_g4": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=4), 16, 2),
    "tiny_cls": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, num_classes=10), 16, 2),
}


WG_CFGS = {
    # WaveGradUNet fixtures (tests/golden/make_golden_wavegrad_unet.py): name -> (cfg, image size, batch)
    "wg_tiny": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, film=True), 16, 2),
    "wg_cfg": (dict(dim=128, dim_mults=[1, 2, 2, 2], channels=3, groups=8, film=True), 32, 1),
    "wg_cls": (dict(dim=32, dim_mults=[1, 2], channels=3, groups=8, film=True, num_classes=10), 16, 2),
}


def wg_inputs(name):
    """The seeded input of the WaveGradUNet fixture (x is regenerated, the noise level is stored)."""
    cfg, size, b = WG_CFGS[name]
    g = torch.Generator().manual_seed(11)
    return torch.randn(b, cfg["channels"], size, size, generator=g)


def make_wavegrad_unet(cfg, sd=None, dtype="fp32", engine="simt", device=None):
    import diffusion_model_nemo_b200.modules as M

    u = M.WaveGradUNet(None, dim=cfg["dim"], dim_mults=cfg["dim_mults"], channels=cfg["channels"], use_convnext=False,
                       resnet_block_groups=cfg["groups"], num_classes=cfg.get("num_classes"), compute_dtype=dtype, conv_engine=engine)
    if sd is not None:
        u.load_state_dict(sd, strict=True)
    if device is not None:
        u = u.to(device)
    return u


def make_unet(cfg, sd=None, dtype="fp32", engine="simt", device=None):
    import diffusion_model_nemo_b200.modules as M

    u = M.Unet(None, dim=cfg["dim"], dim_mults=cfg["dim_mults"], channels=cfg["channels"], use_convnext=False,
               resnet_block_groups=cfg["groups"], learned_variance=cfg.get("learned_variance", False),
               num_classes=cfg.get("num_classes"), compute_dtype=dtype, conv_engine=engine)
    if sd is not None:
        u.load_state_dict(sd, strict=True)
    if device is not None:
        u = u.to(device)
    return u
