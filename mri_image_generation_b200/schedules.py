"""Noise schedules and the registered buffers of the diffusion wrappers (host logic, fp32 on
the CPU exactly as the reference builds them, then moved with the module).

linear:  slice_cond_2d_ddpm/diffusion.py:23-49, ddpm_25d_all_modalities/diffusion.py:22-47
cosine:  ddpm_3d_ldm/diffusion.py:23-56
These must stay bit-identical to the reference (checkpoints carry them; metrics.py:291-294
infers `timesteps` from betas.numel()), which tests/test_oracle_golden.py and tests/test_host_api.py check against
tests/golden/schedules.pt.
"""
import math
from collections import OrderedDict

import torch


def linear_beta_schedule(timesteps: int, beta_start: float, beta_end: float) -> torch.Tensor:
    return torch.linspace(beta_start, beta_end, timesteps, dtype=torch.float32)


def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> torch.Tensor:
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float32)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clamp(betas, 1e-8, 0.999)


def make_buffers(betas: torch.Tensor, with_snr: bool) -> "OrderedDict[str, torch.Tensor]":
    """Buffers in the reference's registration order (state_dict key order)."""
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    alphas_cumprod_prev = torch.cat(
        [torch.tensor([1.0], dtype=torch.float32), alphas_cumprod[:-1]], dim=0)
    out = OrderedDict()
    out["betas"] = betas
    out["alphas"] = alphas
    out["alphas_cumprod"] = alphas_cumprod
    out["alphas_cumprod_prev"] = alphas_cumprod_prev
    out["sqrt_alphas_cumprod"] = torch.sqrt(alphas_cumprod)
    out["sqrt_one_minus_alphas_cumprod"] = torch.sqrt(1.0 - alphas_cumprod)
    out["sqrt_recip_alphas"] = torch.sqrt(1.0 / alphas)
    if with_snr:
        out["snr"] = alphas_cumprod / (1.0 - alphas_cumprod)
    posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
    out["posterior_variance"] = posterior_variance
    out["posterior_log_variance_clipped"] = torch.log(torch.clamp(posterior_variance, min=1e-20))
    return out
