"""Noise schedules and the sampler base class (drop-in for reference modules/diffusion_process.py).

Schedule tables stay plain CPU fp32 tensors, exactly like the reference (its `state_dict()` is empty), and are
produced by the SAME torch CPU op sequence so they are bit-identical; the device only ever sees per-step
coefficient rows derived from them (see `_runtime.py`).
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def cosine_beta_schedule(timesteps, s=0.008, min_clip=0.0001, max_clip=0.999):
    # op order of reference modules/diffusion_process.py:12-17 (https://arxiv.org/abs/2102.09672)
    grid = torch.linspace(0, timesteps, timesteps + 1)
    abar = torch.cos(((grid / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    abar = abar / abar[0]
    return torch.clip(1 - (abar[1:] / abar[:-1]), min_clip, max_clip)


def linear_beta_schedule(timesteps, beta_start=0.0001, beta_end=0.02):
    return torch.linspace(beta_start, beta_end, timesteps)


def quadratic_beta_schedule(timesteps, beta_start=0.0001, beta_end=0.02):
    return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps) ** 2


def sigmoid_beta_schedule(timesteps, beta_start=0.0001, beta_end=0.02):
    return torch.sigmoid(torch.linspace(-6, 6, timesteps)) * (beta_end - beta_start) + beta_start


SCHEDULE_FNS = {
    "linear": linear_beta_schedule,
    "quadratic": quadratic_beta_schedule,
    "sigmoid": sigmoid_beta_schedule,
    "cosine": cosine_beta_schedule,
}


def gaussian_tables(betas: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The 13 derived tables of GaussianDiffusion.compute_constants (reference gaussian_diffusion.py:60-83),
    same expressions in the same order (fp32 CPU)."""
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
        "sqrt_alphas_cumprod": torch.sqrt(abar),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - abar),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / abar),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / abar - 1),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - abar),
        "posterior_variance": post_var,
        # index 0 := index 1 because the posterior variance is 0 at t = 0
        "posterior_log_variance_clipped": torch.log(torch.cat([post_var[1].unsqueeze(0), post_var[1:]])),
        "posterior_mean_coef1": betas * torch.sqrt(abar_prev) / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * torch.sqrt(alphas) / (1.0 - abar),
    }


class AbstractDiffusionProcess(ABC, torch.nn.Module):
    """Common sampler interface (reference modules/diffusion_process.py:39-91)."""

    use_class_conditioning: bool = False

    def __init__(self, timesteps, schedule_name, schedule_cfg=None):
        super().__init__()
        self.timesteps = timesteps
        self.schedule_name = schedule_name
        self.schedule_cfg = schedule_cfg if schedule_cfg is not None else {}
        self.schedule_fn = None

    @abstractmethod
    def compute_constants(self, timesteps):
        raise NotImplementedError()

    @abstractmethod
    def q_mean_variance(self, x_start, t):
        raise NotImplementedError()

    @abstractmethod
    def q_posterior(self, x_start, x, t):
        raise NotImplementedError()

    @abstractmethod
    def q_sample(self, x_start, t, noise=None):
        raise NotImplementedError()

    @abstractmethod
    def p_mean_variance(self, model, x, t, model_output=None):
        raise NotImplementedError()

    @abstractmethod
    def p_sample(self, model, x, t):
        raise NotImplementedError()

    @abstractmethod
    def sample(self, model, shape: List[int], device: torch.device = None):
        raise NotImplementedError()

    def interpolate(self, model, x1, x2, t: Optional[int] = None, lambd: float = 0.0):
        raise NotImplementedError()

    def extract(self, a: torch.Tensor, t: torch.Tensor, x_shape):
        """a[t] broadcast to x's rank: gather on the CPU table, result on t's device."""
        vals = a.gather(-1, t.cpu())
        return vals.reshape(t.shape[0], *((1,) * (len(x_shape) - 1))).to(t.device)

    def forward(self, *args, **kwargs):
        raise RuntimeWarning(f"{self.__class__.__name__} should not be used with forward(), please explicitly call "
                             f"the methods of this module.")


@dataclass
class CosineSchedule:
    s: float = 0.008
    min_clip: float = 0.0001
    max_clip: float = 0.999


@dataclass
class LinearSchedule:
    beta_start: float = 0.0001
    beta_end: float = 0.02


@dataclass
class QuadraticSchedule:
    beta_start: float = 0.0001
    beta_end: float = 0.02


@dataclass
class SigmoidSchedule:
    beta_start: float = 0.0001
    beta_end: float = 0.02
