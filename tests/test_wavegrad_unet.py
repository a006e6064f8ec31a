"""WaveGradUNet (reference modules/unet.py:171-266, parts/film.py): the CPU oracle against the fixture
produced by EXECUTING the reference (tests/golden/make_golden_wavegrad_unet.py), the drop-in's parameter tree, and -- on the
GPU -- the native engine against both."""
import numpy as np
import pytest
import torch

from conftest import WG_CFGS, make_wavegrad_unet, rel_l2, wg_inputs
from oracle import ref_port as O

DEV = "cuda:0"
TOL = {"fp32": 1e-4, "bf16": 2e-2}


@pytest.mark.parametrize("name", list(WG_CFGS))
def test_oracle_matches_reference_golden(golden, name):
    cfg, size, b = WG_CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    x = wg_inputs(name)
    level = torch.from_numpy(golden["wavegrad_unet"][f"{name}/level"])
    cls = torch.tensor([3, 10][:b]) if cfg.get("num_classes") is not None else None
    y = O.wavegrad_unet_forward(sd, cfg, x, level, cls)
    ref = torch.from_numpy(golden["wavegrad_unet"][f"{name}/eps"])
    assert rel_l2(y, ref) <= 1e-6


def test_film_positional_encoding_shape_and_values():
    lv = torch.tensor([0.25, 0.9]).view(2, 1, 1, 1)
    pe = O.film_positional_encoding(lv, 8)
    assert pe.shape == (2, 8, 1, 1)
    e = 1e-4 ** (torch.arange(4, dtype=torch.float32) / 4.0)
    want = torch.cat([(5000 * 0.25 * e).sin(), (5000 * 0.25 * e).cos()])
    assert torch.allclose(pe[0].flatten(), want, atol=1e-6)


def test_dropin_parameter_tree_matches_reference_names():
    """state_dict keys / shapes are the reference's (the fixture generator loads the same dict into the reference strict=True)."""
    cfg, _, _ = WG_CFGS["wg_tiny"]
    u = make_wavegrad_unet(cfg)
    want = O.unet_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in u.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert not any(k.startswith("time_mlp") or ".mlp." in k for k in got)
    # 1 + n_levels + (n_levels - 1) FiLM layers, as the reference constructs them (unet.py:204-210)
    assert len(u.films) == 1 + 2 + 1
    cfg_c, _, _ = WG_CFGS["wg_cls"]
    got_c = {k: tuple(v.shape) for k, v in make_wavegrad_unet(cfg_c).state_dict().items()}
    assert got_c == {k: tuple(v) for k, v in O.unet_param_shapes(cfg_c).items()} and "class_embed.weight" in got_c


def test_dropin_rejects_cpu_and_convnext():
    import diffusion_model_nemo_b200.modules as M
    from diffusion_model_nemo_b200 import _lib as L

    cfg, size, b = WG_CFGS["wg_tiny"]
    u = make_wavegrad_unet(cfg)
    with pytest.raises(L.DmnError):
        u(torch.zeros(b, 3, size, size), torch.full((b, 1, 1, 1), 0.5))
    with pytest.raises(NotImplementedError):
        M.WaveGradUNet(None, dim=32, dim_mults=[1, 2])          # use_convnext defaults to True, as in the reference


def test_wavegrad_loop_oracle_matches_reference_golden(golden):
    cfg, size, b = WG_CFGS["wg_tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    mine = O.sample_wavegrad(O.make_wavegrad_model(sd, cfg), [b, 3, size, size], O.ddpm_tables(12, "linear"), O.NoiseQueue(5))
    ref = torch.from_numpy(golden["wavegrad_unet"]["wg_tiny/loop/linear/12/final01"]) * 2 - 1
    assert float((mine - ref).abs().max()) <= 2e-5


# ---- GPU parity ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(WG_CFGS))
@pytest.mark.parametrize("dtype,engine", [("fp32", "simt"), ("bf16", "simt"), ("bf16", "tcgen05")])
def test_native_wavegrad_unet_vs_reference_golden(golden, name, dtype, engine):
    cfg, size, b = WG_CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_wavegrad_unet(cfg, sd, dtype=dtype, engine=engine, device=DEV)
    x = wg_inputs(name).to(DEV)
    level = torch.from_numpy(golden["wavegrad_unet"][f"{name}/level"]).to(DEV)
    cls = torch.tensor([3, 10][:b], device=DEV) if cfg.get("num_classes") is not None else None
    y = u(x, level, cls)
    ref = torch.from_numpy(golden["wavegrad_unet"][f"{name}/eps"])
    assert y.shape == ref.shape and torch.isfinite(y).all()
    assert rel_l2(y.cpu(), ref) <= TOL[dtype], (name, dtype, engine)
    if cls is not None:       # classes=None: the padding row for every sample; the embedding is added AFTER FiLM 0 read the stem
        ref_none = torch.from_numpy(golden["wavegrad_unet"][f"{name}/eps_noclass"])
        assert rel_l2(u(x, level).cpu(), ref_none) <= TOL[dtype]
        assert rel_l2(ref, ref_none) > 1e-3
    if name == "wg_cfg" and engine == "tcgen05":
        ops = u.plan(size, b, DEV).op_table()
        film_convs = [o for o in ops if o[0].startswith("films.")]
        assert len(film_convs) == 12 and all(o[2] == 1 for o in film_convs)      # every FiLM conv on the tensor-core engine
        assert sum(1 for o in ops if o[1] == 7) == 4                              # one modulation per evaluated FiLM layer


@pytest.mark.gpu
def test_native_wavegrad_unet_per_sample_levels_and_batch(golden):
    """Every sample carries its own noise level; a larger plan serves a smaller batch."""
    cfg, size, b = WG_CFGS["wg_tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_wavegrad_unet(cfg, sd, dtype="fp32", engine="simt", device=DEV)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 3, size, size, generator=g)
    level = torch.rand(5, 1, 1, 1, generator=g) * 0.9 + 0.05
    y = u(x.to(DEV), level.to(DEV))
    ref = O.wavegrad_unet_forward(sd, cfg, x, level)
    assert rel_l2(y.cpu(), ref) <= 1e-4
    y2 = u(x[:2].to(DEV), level[:2].view(-1).to(DEV))            # [B] noise levels are accepted too
    assert rel_l2(y2.cpu(), ref[:2]) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,engine,tol", [("fp32", "simt", 2e-3), ("bf16", "tcgen05", 0.25)])
def test_native_wavegrad_loop_vs_reference_golden(golden, dtype, engine, tol):
    """WaveGradDiffusion.sample(WaveGradUNet) natively (CUDA-graph loop, FiLM encodings tabulated per step) with injected noise
    against the executed reference's final sample."""
    import diffusion_model_nemo_b200.modules as M

    cfg, size, b = WG_CFGS["wg_tiny"]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_wavegrad_unet(cfg, sd, dtype=dtype, engine=engine, device=DEV)
    T = 12
    q = O.NoiseQueue(5)
    noise = torch.stack([q((b, 3, size, size)) for _ in range(T + 1)])
    dif = M.WaveGradDiffusion(timesteps=T, schedule_name="linear")
    imgs = dif.sample(u, [b, 3, size, size], device=DEV, noise=noise)
    ref = torch.from_numpy(golden["wavegrad_unet"]["wg_tiny/loop/linear/12/final01"])
    assert imgs[-1].shape == ref.shape
    assert float((imgs[-1] - ref).abs().max()) <= tol
    # graph replay == plain launches, bit for bit
    dif.use_cuda_graph = False
    imgs2 = dif.sample(u, [b, 3, size, size], device=DEV, noise=noise)
    assert torch.equal(imgs[-1], imgs2[-1])


@pytest.mark.gpu
def test_native_wavegrad_full_batch_properties():
    """BASELINE config 5b shape at full batch (256 x 3 x 32 x 32): finite, deterministic, and the modulation is live (the output
    depends on the noise level)."""
    cfg, size, _ = WG_CFGS["wg_cfg"]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_wavegrad_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(256, 3, size, size, generator=g).to(DEV)
    la, lb = torch.full((256,), 0.3, device=DEV), torch.full((256,), 0.8, device=DEV)
    ya, ya2, yb = u(x, la), u(x, la), u(x, lb)
    assert torch.isfinite(ya).all() and torch.equal(ya, ya2)
    assert rel_l2(ya.cpu(), yb.cpu()) > 1e-3
    # sample 0 of the big batch == the same sample evaluated alone against the oracle tolerance
    ref = O.wavegrad_unet_forward(sd, cfg, x[:1].cpu(), la[:1].view(1, 1, 1, 1).cpu())
    assert rel_l2(ya[:1].cpu(), ref) <= 2e-2
