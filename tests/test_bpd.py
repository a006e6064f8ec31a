"""Bits-per-dimension evaluation (reference models/abstract_diffusion_model.py:137-197,
loss/variational_bound_loss.py:31-52, utils.py:28-56): the CPU oracle against the fixture produced by EXECUTING the reference
(tests/golden/make_golden_bpd.py), and -- on the GPU -- the native loop (q_sample kernel, U-Net, fused term reduction) against both."""
import pytest
import torch

from conftest import CFGS, make_unet
from oracle import ref_port as O

DEV = "cuda:0"
CASES = {
    # key prefix in bpd.npz: (unet cfg name, schedule, T, learned variance, q_sample noise seed)
    "ddpm/linear/20": ("tiny", "linear", 20, False, 7),
    "learned/cosine/16": ("tiny_lv", "cosine", 16, True, 9),
}


def _noise(seed, T, shape):
    q = O.NoiseQueue(seed)
    return torch.stack([q(tuple(shape)) for _ in range(T)])


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_golden(golden, case):
    name, sched, T, learned, seed = CASES[case]
    cfg, _, _ = CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    x0 = torch.from_numpy(golden["bpd"][f"{case}/x0"])
    mine = O.bits_per_dimension(O.make_model(sd, cfg), x0, O.ddpm_tables(T, sched), O.NoiseQueue(seed), learned=learned)
    for k in ("total_bpd", "terms_bpd", "prior_bpd"):
        ref = torch.from_numpy(golden["bpd"][f"{case}/{k}"])
        assert mine[k].shape == ref.shape
        assert torch.allclose(mine[k], ref, rtol=1e-5, atol=1e-6), k


def test_decoder_edge_bins_and_kl_identities():
    """utils.discretized_gaussian_log_likelihood: the open-ended first / last bins (|x| > 0.999); normal_kl(p, p) = 0."""
    x = torch.tensor([-1.0, 1.0, 0.0])
    m = torch.zeros(3)
    ll = O.discretized_gaussian_log_likelihood(x, m, torch.zeros(3))
    assert torch.allclose(ll[0], ll[1], rtol=1e-5) and ll[2] < 0 and torch.isfinite(ll).all()
    z = torch.randn(5)
    assert torch.allclose(O.normal_kl(z, z * 0.3, z, z * 0.3), torch.zeros(5), atol=1e-7)


def test_sampler_rejects_cpu():
    import diffusion_model_nemo_b200.modules as M
    from diffusion_model_nemo_b200 import _lib as L

    s = M.GaussianDiffusion(10, "linear")
    with pytest.raises(L.DmnError):
        s.calculate_bits_per_dimension(torch.zeros(2, 1, 16, 16), lambda x, t: x)


# ---- GPU parity ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("use_graph", [True, False])
def test_native_bpd_loop_fp32_vs_reference_golden(golden, case, use_graph):
    import diffusion_model_nemo_b200.modules as M

    name, sched, T, learned, seed = CASES[case]
    cfg, _, _ = CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="fp32", engine="simt", device=DEV)
    x0 = torch.from_numpy(golden["bpd"][f"{case}/x0"]).to(DEV)
    s = (M.LearnedGaussianDiffusion if learned else M.GaussianDiffusion)(T, sched)
    s.use_cuda_graph = use_graph
    out = s.calculate_bits_per_dimension(x0, u, noise=_noise(seed, T, x0.shape))
    for k, rtol in (("prior_bpd", 1e-5), ("terms_bpd", 5e-3), ("total_bpd", 5e-3)):
        ref = torch.from_numpy(golden["bpd"][f"{case}/{k}"])
        got = out[k].cpu()
        assert got.shape == ref.shape and torch.isfinite(got).all()
        assert torch.allclose(got, ref, rtol=rtol, atol=1e-5), (k, float((got - ref).abs().max()))


@pytest.mark.gpu
def test_native_bpd_foreign_model_and_max_batch(golden):
    """Any (x, t) callable works (evaluated per step, kernels still native); max_batch_size truncates the batch as the reference does."""
    import diffusion_model_nemo_b200.modules as M

    name, sched, T, learned, seed = CASES["ddpm/linear/20"]
    cfg, _, _ = CFGS[name]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="fp32", engine="simt", device=DEV)
    x0 = torch.from_numpy(golden["bpd"]["ddpm/linear/20/x0"]).to(DEV)
    s = M.GaussianDiffusion(T, sched)
    noise = _noise(seed, T, x0.shape)
    a = s.calculate_bits_per_dimension(x0, u, noise=noise)
    b = s.calculate_bits_per_dimension(x0, lambda x, t: u(x, t), noise=noise)
    assert torch.allclose(a["terms_bpd"], b["terms_bpd"], rtol=1e-5, atol=1e-6)
    one = s.calculate_bits_per_dimension(x0, u, max_batch_size=1, noise=noise[:, :1])
    assert one["terms_bpd"].shape == (1, T) and torch.allclose(one["terms_bpd"], a["terms_bpd"][:1], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_native_bpd_bf16_tensor_core_and_philox():
    """Throughput mode on the CIFAR-shape U-Net (bf16 / tcgen05, in-kernel Philox q_sample noise): finite, deterministic per seed,
    every term non-negative up to round-off (KL >= 0, NLL >= 0) and the prior term independent of the network."""
    import diffusion_model_nemo_b200.modules as M

    cfg, size, _ = CFGS["cfg2"]
    sd = O.random_state_dict(cfg, seed=0)
    u = make_unet(cfg, sd, dtype="bf16", engine="tcgen05", device=DEV)
    g = torch.Generator().manual_seed(5)
    x0 = (torch.rand(8, 3, size, size, generator=g) * 2 - 1).to(DEV)
    s = M.GaussianDiffusion(50, "linear")
    s.seed = 11
    a = s.calculate_bits_per_dimension(x0, u)
    b = s.calculate_bits_per_dimension(x0, u)
    assert torch.equal(a["terms_bpd"], b["terms_bpd"]) and torch.isfinite(a["total_bpd"]).all()
    assert float(a["terms_bpd"].min()) > -1e-4
    tb = O.ddpm_tables(50, "linear")
    prior = O.bits_per_dimension(lambda x, t: torch.zeros_like(x), x0.cpu(), tb, O.NoiseQueue(0))["prior_bpd"]
    assert torch.allclose(a["prior_bpd"].cpu(), prior, rtol=1e-5, atol=1e-7)
